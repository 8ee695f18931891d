"""GPU probe: per-op CUDA-event time of one eager video-mode forward (T=10, 256x256)."""
import sys, collections, types
import torch
sys.path.insert(0, ".")
from flair_b200 import ops, synth
import guided_diffusion.unet_new as U
from guided_diffusion.script_util import blur_unet_config

T = int(sys.argv[1]) if len(sys.argv) > 1 else 10
model = U.UNetModel(**blur_unet_config(256))
model.load_state_dict(synth.synthetic_state_dict(model)); model.convert_to_fp16(); model.eval().cuda()
model.use_cuda_graph = False
dev = "cuda"
clip = (synth.synthetic_clip(T, 256) * 2 - 1).to(dev); x = torch.randn(T, 3, 256, 256, device=dev); ts = torch.full((T,), 500, device=dev)
fwd = lambda: model(x, ts, low_res_input=clip[None], num_frames=T, enable_cross_frames=True, vsrpp_weights=1.0)
fwd(); fwd()
rec = []
def wrap(name, fn):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); r = fn(*a, **k); e1.record()
        extra = ""
        if name == "conv":
            xx = a[0]; extra = f" {tuple(xx.shape[1:4])} {xx.shape[-1]}->{a[2]} k{k.get('ksize', a[3] if len(a) > 3 else (1,3,3))}"
        rec.append((name + extra, e0, e1)); return r
    return w
names = ["conv", "gn_stats", "gn_apply", "concat_channels", "copy_channels_into", "attn_spatial", "attn_temporal",
         "linear_f32", "pack_im2col6", "flow_warp", "flow_compose", "planes_to_cl", "deform_im2col", "scale_pixels_", "timestep_embedding"]
for n in names: setattr(ops, n, wrap(n, getattr(ops, n)))
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record(); fwd(); e1.record(); torch.cuda.synchronize()
tot = collections.Counter(); cnt = collections.Counter()
for n, a, b in rec:
    tot[n] += a.elapsed_time(b); cnt[n] += 1
print("eager forward", e0.elapsed_time(e1), "ms; sum of op events", sum(tot.values()))
agg = collections.Counter(); aggc = collections.Counter()
for n in tot: agg[n.split(" ")[0]] += tot[n]; aggc[n.split(" ")[0]] += cnt[n]
for n, v in agg.most_common(): print(f"{v:8.2f} ms {aggc[n]:5d}  {n}")
print("--- top conv shapes")
for n, v in [kv for kv in tot.most_common() if kv[0].startswith("conv")][:22]: print(f"{v:8.2f} ms {cnt[n]:5d}  {n}  ({v/cnt[n]*1e3:.1f} us each)")
