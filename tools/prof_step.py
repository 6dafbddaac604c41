import sys, time, cProfile, pstats, io
sys.path.insert(0, '/root/repo')
import torch
import bench
from spurfies_b200.train import TrainStep
dev = torch.device('cuda', 0)
sc, model = bench.build_scene(dev, precision='bf16')
step = TrainStep(model)
hb = bench.host_batches(4, 0); db = [bench.to_device(h, dev) for h in hb]
def one(i):
    b, g, r = bench.split(db[i % 4]); return step(b, g, r)
for i in range(4): one(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(10): one(i)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"10 steps: host enqueue {1e3*(t1-t0)/10:.2f} ms/step, total {1e3*(t2-t0)/10:.2f} ms/step, mem {torch.cuda.max_memory_allocated()/1e9:.2f} GB reserved {torch.cuda.memory_reserved()/1e9:.2f} GB")
pr = cProfile.Profile(); pr.enable()
for i in range(5): one(i)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35); print(s.getvalue()[:6000])
