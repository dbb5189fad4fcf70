"""Per-call device-time table of one training step (CUDA events around every C-ABI call, side streams off).
    python scripts/profile_step.py [out.json] [--workload c2|c3|c4|c5]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from genesis_b200 import ops, profiling, trainer
ops.set_side_streams(False)
_argv = sys.argv[1:]
if '--workload' in _argv:
    _i = _argv.index('--workload')
    bench.select_workload(_argv[_i + 1])
    _argv = _argv[:_i] + _argv[_i + 2:]

plugin, cfg = bench.build_cfg()
torch.manual_seed(0)
dev = torch.device('cuda', 0)
model = plugin.load(cfg).to(dev).train()
ts = trainer.TrainStep(model, world_size=1, img_size=bench.IMG)
x = bench.synthetic_batches(1, bench.B_PER_GPU, 7)[0].to(dev)
for _ in range(3):
    ts.step_device(x)
torch.cuda.synchronize()
with profiling.Profiler() as prof:
    ts.step_device(x)
rows = prof.table(by_shape=True)
tot = sum(r['ms'] for r in rows)
print('total device ms in C-ABI calls: %.3f' % tot)
out = []
agg = {}
for r in rows:
    a = agg.setdefault(r['key'][0], [0.0, 0])
    a[0] += r['ms']
    a[1] += r['calls']
print('-- by entry point')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%8.3f ms %4d  %s' % (v[0], v[1], k))
print('-- by shape')
for r in rows[:70]:
    name, shape = r['key']
    tf = r['flops'] / (r['ms'] * 1e-3) / 1e12 if r['flops'] else 0
    gb = r['bytes'] / (r['ms'] * 1e-3) / 1e9 if r['bytes'] else 0
    print('%8.3f ms %3d  %-24s %7.1f TF/s %7.0f GB/s  %s' % (r['ms'], r['calls'], name, tf, gb, shape))
    out.append(dict(ms=r['ms'], calls=r['calls'], name=name, tflops=tf, gbs=gb, shape=list(shape)))
json.dump(out, open(_argv[0] if _argv else 'gpurun_out/step_profile.json', 'w'), indent=0)
