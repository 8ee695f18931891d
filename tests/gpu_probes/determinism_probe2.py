"""GPU probe: find the first op call whose output differs between two identical eager forwards."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from flair_b200 import ops, synth
import guided_diffusion.unet_new as U
S, T = 64, 5
cfg = dict(image_size=S, in_channels=6, model_channels=128, out_channels=6, num_res_blocks=1,
           attention_resolutions=(4,), rnn_resolutions=(1, 2), channel_mult=(0.5, 1, 4), num_head_channels=64,
           resblock_updown=True, use_scale_shift_norm=True, temporal_block=True, use_fp16=True)
model = U.UNetModel(**cfg)
model.load_state_dict(synth.synthetic_state_dict(model, seed=99)); model.convert_to_fp16(); model.eval().cuda()
model.use_cuda_graph = False
dev = "cuda"
clip = (synth.synthetic_clip(T, S) * 2 - 1).to(dev)
ts = torch.full((T,), 500, device=dev)
x = torch.randn(T, 3, S, S, device=dev)
fwd = lambda: model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
fwd()
log = []
def wrap(name, fn):
    def w(*a, **k):
        r = fn(*a, **k)
        ins = [t.clone() for t in a if torch.is_tensor(t)]
        outs = r if isinstance(r, (tuple, list)) else [r]
        log.append((name, ins, [t.clone() for t in outs if torch.is_tensor(t)],
                    [k[n].clone() for n in ("out2",) if k.get(n) is not None]))
        return r
    return w
for n in ["conv", "gn_stats", "gn_apply", "deform_conv", "flow_warp", "copy_channels_into", "planes_to_cl", "attn_spatial", "attn_temporal"]:
    setattr(ops, n, wrap(n, getattr(ops, n)))
fwd(); A = log; log = []
fwd(); B = log
torch.cuda.synchronize()
print("calls", len(A), len(B))
for i, (a, b) in enumerate(zip(A, B)):
    same_in = all(torch.equal(p, q) for p, q in zip(a[1], b[1]))
    same_out = all(torch.equal(p, q) for p, q in zip(a[2], b[2]))
    same_o2 = all(torch.equal(p, q) for p, q in zip(a[3], b[3]))
    if not (same_out and same_o2):
        print(f"call {i} {a[0]}: inputs same {same_in}, outputs same {same_out}, out2 same {same_o2}; input shapes {[tuple(t.shape) for t in a[1]]}")
        if not same_in:
            for j, (p, q) in enumerate(zip(a[1], b[1])):
                if not torch.equal(p, q):
                    d = (p.float() - q.float()).abs()
                    print(f"   input {j} shape {tuple(p.shape)} differing {int((d > 0).sum())} max {float(d.max()):.3e}")
        break
else:
    print("no differing call")
