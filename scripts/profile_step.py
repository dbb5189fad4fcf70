"""Per-call device-time table of one GENESIS training step (CUDA events around every C-ABI call)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from genesis_b200 import ops, profiling, trainer
ops.set_side_streams(False)

plugin, cfg = bench.build_cfg()
torch.manual_seed(0)
dev = torch.device('cuda', 0)
model = plugin.load(cfg).to(dev).train()
ts = trainer.TrainStep(model, world_size=1)
x = torch.rand(bench.B_PER_GPU, 3, 64, 64, device=dev)
for _ in range(3):
    ts.step_device(x)
torch.cuda.synchronize()
with profiling.Profiler() as prof:
    ts.step_device(x)
rows = prof.table(by_shape=True)
tot = sum(r['ms'] for r in rows)
print('total device ms in C-ABI calls: %.3f' % tot)
out = []
for r in rows[:60]:
    name, shape = r['key']
    tf = r['flops'] / (r['ms'] * 1e-3) / 1e12 if r['flops'] else 0
    gb = r['bytes'] / (r['ms'] * 1e-3) / 1e9 if r['bytes'] else 0
    print('%8.3f ms %3d  %-24s %7.1f TF/s %7.0f GB/s  %s' % (r['ms'], r['calls'], name, tf, gb, shape))
    out.append(dict(ms=r['ms'], calls=r['calls'], name=name, tflops=tf, gbs=gb, shape=list(shape)))
json.dump(out, open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/step_profile.json', 'w'), indent=0)
