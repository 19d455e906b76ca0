import os, sys, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + '/tests')
import bench
from torch.profiler import profile, ProfilerActivity
cfg = dict(bench.CFG, workload='full')
dev = torch.device('cuda:0')
torch.backends.cudnn.benchmark = True
pipe, params = bench.build_pipeline(dev)
flat = {k: v.to(dev) for k, v in bench.flatten_batch(bench.make_full_batch(cfg, 0)).items()}
def step():
    for p in params: p.grad = None
    out = pipe(bench.unflatten_batch(flat), 0, discriminate=True)
    bench.full_loss(out, flat['target.image']).backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True, with_stack=False) as prof:
    step(); torch.cuda.synchronize()
ka = prof.key_averages(group_by_input_shape=True)
rows = [e for e in ka if e.self_device_time_total > 0 and (e.key.startswith('aten::') or e.key.startswith('_'))]
rows.sort(key=lambda e: -e.self_device_time_total)
tot = sum(e.self_device_time_total for e in rows)
print(f"total self device time {tot/1e3:.1f} ms")
for e in rows[:150]:
    print(f"{e.self_device_time_total/1e3:8.2f} ms {e.count:4d}x {e.key[:40]:40s} {str(e.input_shapes)[:150]}")
