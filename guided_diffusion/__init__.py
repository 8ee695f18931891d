"""Drop-in Python boundary of the B200-native FLAIR hot path.

Same module paths and public names as the reference's `guided_diffusion`
package, but every hot-path body launches hand-written sm_100a
kernels from `flair_b200` through its C ABI.  There is no CPU/PyTorch fallback:
tensors must live on a B200.
"""
