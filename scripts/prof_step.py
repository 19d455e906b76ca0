import sys, cProfile, pstats, torch, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,R); sys.path.insert(0,R+'/tests')
import bench
cfg=dict(bench.CFG)
from latentsplat_b200.model.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
dev=torch.device('cuda:0')
dec=DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda"),[0.,0.,0.]).to(dev)
batch=bench.make_batch(cfg,0)
db={k:v.to(dev) for k,v in batch.items()}
leaves={k:db[k].clone().requires_grad_(True) for k in bench.GAUSSIAN_KEYS}
for _ in range(5): bench.step_device(dec,db,leaves)
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(20): bench.step_device(dec,db,leaves)
torch.cuda.synchronize()
print("ms/step", (time.perf_counter()-t0)*50)
pr=cProfile.Profile(); pr.enable()
for _ in range(20): bench.step_device(dec,db,leaves)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
