import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + '/tests')
import bench
from torch.profiler import profile, ProfilerActivity, record_function
cfg = dict(bench.CFG, workload='full')
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
pipe, params = bench.build_pipeline(dev)
pipe.encoder.backbone.dino = pipe.encoder.backbone.dino if hasattr(pipe.encoder.backbone, "dino") else None
flat = {k: v.to(dev) for k, v in bench.flatten_batch(bench.make_full_batch(cfg, 0)).items()}
if len(sys.argv) > 1 and sys.argv[1] == 'tf32':
    torch.backends.cuda.matmul.allow_tf32 = True
def step():
    for p in params: p.grad = None
    with record_function("FWD"):
        out = pipe(bench.unflatten_batch(flat), 0, discriminate=True)
        loss = bench.full_loss(out, flat['target.image'])
    with record_function("BWD"):
        loss.backward()
for _ in range(2): step()
torch.cuda.synchronize()
# coarse module timing with hooks
times = {}
def hook(name):
    def pre(m, i):
        e = torch.cuda.Event(enable_timing=True); e.record(); m._t0 = e
    def post(m, i, o):
        e = torch.cuda.Event(enable_timing=True); e.record(); times.setdefault(name, []).append((m._t0, e))
    return pre, post
mods = {"dino": pipe.encoder.backbone.dino, "backbone_mlps+proj": None, "epipolar_transformer": pipe.encoder.epipolar_transformer,
        "epi.sampler": pipe.encoder.epipolar_transformer.epipolar_sampler, "epi.transformer": pipe.encoder.epipolar_transformer.transformer,
        "epi.refine": pipe.encoder.epipolar_transformer.upscale_refinement, "hi_res_skip": pipe.encoder.high_resolution_skip,
        "gaussian_adapter": pipe.encoder.gaussian_adapter, "depth_predictor": pipe.encoder.depth_predictor,
        "encoder": pipe.encoder, "decoder(splat)": pipe.decoder, "vae.decoder.mid": pipe.autoencoder.model.decoder.mid_block,
        "vae.up0": pipe.autoencoder.model.decoder.up_blocks[0], "vae.up1": pipe.autoencoder.model.decoder.up_blocks[1],
        "vae.up2": pipe.autoencoder.model.decoder.up_blocks[2], "vae.up3": pipe.autoencoder.model.decoder.up_blocks[3]}
hs = []
for n, m in mods.items():
    if m is None: continue
    a, b = hook(n); hs += [m.register_forward_pre_hook(a), m.register_forward_hook(b)]
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
for p in params: p.grad = None
e0.record(); out = pipe(bench.unflatten_batch(flat), 0, discriminate=True); loss = bench.full_loss(out, flat['target.image']); e1.record(); loss.backward(); e2.record()
torch.cuda.synchronize()
print(f"FWD {e0.elapsed_time(e1):.1f} ms  BWD {e1.elapsed_time(e2):.1f} ms")
for n, v in times.items():
    print(f"  fwd {n:24s} {sum(a.elapsed_time(b) for a, b in v):8.2f} ms ({len(v)} calls)")
for h in hs: h.remove()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
ka = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in ka)
print(f"Self CUDA total {tot/1e3:.1f} ms")
print("---- aten ops (self device time) ----")
for e in [e for e in ka if e.key.startswith('aten::')][:45]:
    print(f"{e.self_device_time_total/1e3:9.2f} ms {100*e.self_device_time_total/tot:5.1f}% {e.count:5d}x  {e.key[:110]}")
print("---- kernels ----")
ka = [e for e in ka if not e.key.startswith('aten::') and e.key not in ('FWD','BWD') and 'Backward' not in e.key and not e.key.startswith('_') and 'autograd' not in e.key]
for e in ka[:50]:
    print(f"{e.self_device_time_total/1e3:9.2f} ms {100*e.self_device_time_total/tot:5.1f}% {e.count:5d}x  {e.key[:110]}")

# ---- where does the remaining torch glue come from: aten ops by (op, input shapes, first repo frame) ----
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=True) as prof2:
    step(); torch.cuda.synchronize()
import collections
agg = collections.OrderedDict()
for ev in prof2.events():
    if not ev.name.startswith("aten::") or ev.self_device_time_total <= 0:
        continue
    frame = next((f for f in (ev.stack or []) if "latentsplat_b200" in f or "bench.py" in f), "?")
    frame = frame.split("latentsplat_b200/")[-1][:70]
    shapes = str([s for s in (ev.input_shapes or []) if s])[:70]
    a = agg.setdefault((ev.name, shapes, frame), [0, 0.0])
    a[0] += 1
    a[1] += ev.self_device_time_total
print("---- aten glue by (op, shapes, repo frame) ----")
for (name, shapes, frame), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t/1e3:8.2f} ms {n:4d}x {name[6:]:28s} {shapes:70s} {frame}")
