"""Device-side part of the reference's facelib that sits on the sampler's hot path: the affine crops
and inverse warps of the aux face prior.  Detection (RetinaFace), parsing (ParseNet) and the prior (CodeFormer) remain
reference PyTorch modules supplied by the caller (BASELINE north_star)."""
