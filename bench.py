#!/usr/bin/env python
"""bench.py -- headline metric of BASELINE.json: images/sec, forward + backward (+ gradient all-reduce and
Adam), GENESIS K=5 on 64x64 synthetic Multi-dSprites-shaped batches, B=64 per GPU, at N GPUs of one node.

    python bench.py --gpus N --steps K --warmup W              # engine (one process per GPU; torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W      # reference arm: the reference's own nn.Modules
                                                               # (oracle/_ref; else the oracle port) on the host cores
    python bench.py --workload c3|c4|c5 ...                    # the other BASELINE configs' per-GPU shapes

One JSON line on stdout (rank 0).  See DESIGN.md section 5 for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'images/sec fwd+bwd 64x64 K=5'
# Headline workload = BASELINE.json configs[1] (c2).  --workload c3/c4/c5 time the other configs' per-GPU shapes
# (extra measurements; the driver only runs the default).
WORKLOADS = {   # name: (model, K, img, batch per GPU, generator, fwd GFLOP / image as the engine computes it: SURVEY.md 8d)
    'c2': ('genesis', 5, 64, 64, 'multid', 6.059),
    'c3': ('genesisv2', 7, 64, 128, 'stacks', 3.675),
    'c4': ('genesisv2', 11, 64, 32, 'rooms', 4.801),
    'c5': ('monet', 7, 128, 64, 'multid', 14.022),
}
MODEL, K_SLOTS, IMG, B_PER_GPU, GEN, FWD_GFLOP_PER_IMG = WORKLOADS['c2']


def workload_name():
    return '%s K=%d %dx%d %s-shaped synthetic, batch %d per GPU' % (MODEL, K_SLOTS, IMG, IMG, GEN, B_PER_GPU)


def select_workload(name):
    global MODEL, K_SLOTS, IMG, B_PER_GPU, GEN, FWD_GFLOP_PER_IMG, METRIC
    MODEL, K_SLOTS, IMG, B_PER_GPU, GEN, FWD_GFLOP_PER_IMG = WORKLOADS[name]
    METRIC = 'images/sec fwd+bwd %dx%d K=%d' % (IMG, IMG, K_SLOTS)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sus=d['bf16_tflops_sustained'], src='measured')
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src='fallback')


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [t.strip() for t in out.strip().split(',')]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unsampled']}
        sm = sorted(int(float(s[0])) for s in self.samples)
        reasons = []
        for i, name in enumerate(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')):
            if any(s[2 + i].lower().startswith('active') for s in self.samples):
                reasons.append(name)
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': int(float(self.samples[0][1])), 'reasons': reasons,
                'samples': len(sm)}


def synthetic_batches(n_batches, batch, seed):
    from genesis_b200.datasets import synth    # the product's own generator of dataset-shaped images
    return [torch.from_numpy(synth.GENERATORS[GEN](batch, IMG, seed + i)[0]) for i in range(n_batches)]


def build_cfg():
    import importlib
    plugin = importlib.import_module('genesis_b200.model_configs.%s_config' % MODEL)
    from forge import flags
    cfg = dict(flags.defaults())
    cfg.update(debug=False, multi_gpu=False, img_size=IMG, K_steps=K_SLOTS)
    from attrdict import AttrDict
    return plugin, AttrDict(cfg)


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(batch):
    """One training step of the reference's PyTorch path on the host: forward + loss assembly (train.py:227-259 without GECO's
    host sync) + backward.  Preferred: the reference's OWN nn.Module from oracle/_ref (byte-identical copy of models/,
    modules/, third_party/sylvester; kind "reference"); fallback: the oracle port (oracle/models.py; kind "port").
    Returns (step function, kind)."""
    from oracle import ref_loader
    xs = synthetic_batches(2, batch, 100)
    if ref_loader.available():
        from oracle import models as M
        cfg = dict(M.make_cfg(MODEL, K_steps=K_SLOTS, img_size=IMG))
        model = ref_loader.load_reference(MODEL, cfg, seed=0).train()

        def step(i):
            model.zero_grad(set_to_none=True)
            torch.manual_seed(i)
            recon, losses, stats, att, comp = model(xs[i % len(xs)])
            err = losses.err.mean(0)
            kl = err.new_zeros(())
            for key in ('kl_m_k', 'kl_l_k'):
                if key in losses and len(losses[key]):
                    kl = kl + torch.stack(list(losses[key]), dim=1).mean(0).sum()
            for key in ('kl_m', 'kl_l'):
                if key in losses and torch.is_tensor(losses[key]):
                    kl = kl + losses[key].mean(0)
            (err + kl).backward()
            return float(err.detach())
        return step, 'reference'
    from oracle import models as M, functional as O
    plugin, cfg = build_cfg()
    torch.manual_seed(0)
    holder = plugin.load(cfg)
    P = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in holder.state_dict().items()}
    ocfg = M.make_cfg(MODEL, K_steps=K_SLOTS, img_size=IMG)

    def step(i):
        for p in P.values():
            if p.is_floating_point() and p.grad is not None:
                p.grad = None
        out = M.FORWARD[MODEL](P, xs[i % len(xs)], O.NoiseTape(seed=i), ocfg, training=True)
        M.total_loss(out).backward()
        return float(out['err'].mean())
    return step, 'port'


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = B_PER_GPU                     # the config's own per-GPU batch (c2: 64), not a reduced sample
    step, kind = cpu_step_fn(batch)
    for i in range(max(1, min(args.warmup, 2))):
        step(i)
    # a bounded sample of the workload: at most args.steps steps and at most ~150 s of host time (never fewer than 2 steps)
    t0 = time.perf_counter()
    n_run = 0
    for i in range(args.steps):
        step(i)
        n_run += 1
        if n_run >= 2 and time.perf_counter() - t0 > 150.0:
            break
    dt = time.perf_counter() - t0
    v = batch * n_run / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * dt / n_run, 'steps_timed': n_run, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': workload_name() + ', fwd+bwd on the host CPU', 'global_batch': batch,
                   'same_config': True},
        'cpu_baseline': {'value': v, 'unit': 'images/s', 'cores': cores, 'kind': kind,
                         'sample': '%d steps x batch %d (the config batch), torch %s CPU fp32, %d threads; %s' % (
                             n_run, batch, torch.__version__, cores,
                             "the reference's own nn.Modules from oracle/_ref" if kind == 'reference' else 'oracle port (oracle/_ref absent)')},
        'e2e': {'value': v, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def run_engine(args):
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the engine has no CPU path')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        import datetime
        dist.init_process_group('nccl', device_id=dev, timeout=datetime.timedelta(seconds=180))

    from genesis_b200 import build, _lib, profiling, trainer
    if rank == 0:
        build.build()
    if world > 1:
        dist.barrier()
    plugin, cfg = build_cfg()
    torch.manual_seed(0)
    model = plugin.load(cfg).to(dev).train()
    # same parameters on every rank (manual_seed(0) above), independent eps / IC-SBP seed noise per rank (noise_seed + rank)
    ts = trainer.TrainStep(model, lr=1e-4, img_size=IMG, world_size=world, rank=rank, noise_seed=1234, overlap=not args.no_overlap)
    lib = _lib.lib()

    n_in = 4
    host = [t.pin_memory() for t in synthetic_batches(n_in, B_PER_GPU, 1000 * rank + 1)]
    devx = [t.to(dev) for t in host]
    graphed = False
    if not args.no_graph:
        try:
            ts.capture(devx[0])
            graphed = True
        except Exception as exc:   # report, do not hide: the eager path is still our kernels, just launch-bound
            sys.stderr.write('CUDA graph capture failed (%s: %s); running eagerly\n' % (type(exc).__name__, exc))
            ts.graph = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- leg 1: device-resident inputs (value)
    for i in range(max(3, args.warmup)):
        ts.step_device(devx[i % n_in])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = lib.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        elbo = ts.step_device(devx[i % n_in])
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = (ts.launches_per_step * args.steps) if graphed else (lib.launches - l0)
    # ---- leg 2: end to end through the public API, pinned host inputs, loss read back every step
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = 0.0
    for i in range(args.steps):
        last = float(ts.step(host[i % n_in]).item())         # H2D copy in, D2H of the ELBO out, every step
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)
    sampler.stop_flag = True
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()

    # ---- dominant kernel: per-call device time inside profiled (eager) steps, CUDA events on the launch stream.
    # Every rank runs them (the step contains the gradient all-reduce); rank 0 reports.
    from genesis_b200 import ops
    ops.set_side_streams(False)       # serial: per-kernel event times are not inflated by concurrent side-stream kernels
    with profiling.Profiler() as prof:
        for i in range(2):
            ts._step_eager(devx[i % n_in])
    ops.set_side_streams(True)
    barrier()
    if rank == 0:
        pk = peaks()
        rows = prof.table()
        total_ms = sum(r['ms'] for r in rows)
        top = rows[0]
        # DRAM bytes per C-ABI CALL of the dominant entry point from the committed ncu pass over one eager step of this
        # workload (profiles/r02_traffic_<workload>.json: dram__bytes_read.sum + dram__bytes_write.sum summed over the kernel
        # launches of that entry point, divided by its calls per step) -- the same unit as the algorithmic bytes / flops
        # below, which are per call too (one call = 1 kernel launch, or 4 for the sub-pixel classes of a stride-2 layer).
        traffic, traffic_src = None, None
        calls_per_step = top['calls'] // 2
        tpath = os.path.join(ROOT, 'profiles', 'r02_traffic_%s.json' % args.workload)
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            if top['key'] in tj and calls_per_step > 0:
                traffic = tj[top['key']]['dram_bytes'] / calls_per_step
                traffic_src = ('profiles/r02_traffic_%s.json: ncu dram bytes of the %d kernel launches of this entry point in one eager '
                               'step / %d calls per step' % (args.workload, tj[top['key']]['launches'], calls_per_step))
        tensor_like = top['flops'] > 0 and top['key'].startswith(('g2_conv', 'g2_gemm'))
        common = {'kernel': top['key'], 'traffic': traffic, 'traffic_source': traffic_src, 'unit_of_launch': 'one C-ABI call',
                  'share_of_step': top['ms'] / total_ms, 'launch_ms': top['ms'] / top['calls'], 'calls_per_step': calls_per_step,
                  'algorithmic_bytes_per_launch': top['bytes'] / top['calls'],
                  'timing': 'CUDA events around each call on the launch stream, 2 eager steps, side streams off'}
        if tensor_like:
            ach = top['flops'] / (top['ms'] * 1e-3) / 1e12
            roof = dict(common, bound='tensor', achieved=ach, peak=pk['tf_sus'], unit='TFLOP/s', frac=ach / pk['tf_sus'],
                        peak_source=pk['src'] + ' bf16 sustained (kernel runs TF32 operands: nominal TF32 dense peak is half of bf16)',
                        algorithmic_flops_per_launch=top['flops'] / top['calls'])
        else:
            ach = top['bytes'] / (top['ms'] * 1e-3) / 1e9
            roof = dict(common, bound='hbm', achieved=ach, peak=pk['hbm'], unit='GB/s', frac=ach / pk['hbm'], peak_source=pk['src'])
        breakdown = [{'kernel': r['key'], 'ms_per_step': r['ms'] / 2, 'calls_per_step': r['calls'] // 2,
                      'tflops': (r['flops'] / (r['ms'] * 1e-3) / 1e12) if r['flops'] else None,
                      'gbs': (r['bytes'] / (r['ms'] * 1e-3) / 1e9) if r['bytes'] else None} for r in rows[:8]]
        step_tf = 3 * FWD_GFLOP_PER_IMG * B_PER_GPU * world * args.steps / (ms * 1e-3) / 1e3
        # ---- CPU baseline: bounded sample of the same workload (the config's own batch) on the host cores
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            step, kind = cpu_step_fn(B_PER_GPU)
            step(0)
            t0 = time.perf_counter()
            n_cpu = 0
            while n_cpu < 3 and (n_cpu == 0 or time.perf_counter() - t0 < 30.0):
                step(n_cpu + 1)
                n_cpu += 1
            dt = time.perf_counter() - t0
            cpu = {'value': B_PER_GPU * n_cpu / dt, 'unit': 'images/s', 'cores': cores, 'kind': kind,
                   'sample': '1 warm-up + %d steps of batch %d (the config batch), %s, torch CPU fp32, %d threads' % (
                       n_cpu, B_PER_GPU, "the reference's own nn.Modules (oracle/_ref)" if kind == 'reference' else 'oracle port', cores)}
        imgs = B_PER_GPU * world * args.steps
        line = {
            'metric': METRIC, 'value': imgs / (ms * 1e-3), 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'tf32', 'data': 'synthetic',
            'config': {'workload': workload_name() + ', fwd+bwd+allreduce+Adam+GECO',
                       'global_batch': B_PER_GPU * world, 'parallelism': 'dp%d' % world,
                       'l2': 'per-step working set (activations > 4 GB) exceeds the 126 MB L2; inputs rotate over %d batches' % n_in,
                       'step_tflops_algorithmic': step_tf, 'last_elbo': last, 'cuda_graph': graphed,
                       'precision': 'tf32 tensor-core operands, fp32 accumulate / storage',
                       'streams': 'one captured graph; prior/KL branch and parameter-gradient kernels on side streams',
                       'allreduce': ('none (1 GPU)' if world == 1 else 'single collective after backward' if args.no_overlap else
                                     '%d gradient buckets all-reduced under the backward pass + 1 tail collective' % len(ts.buckets))},
            'e2e': {'value': imgs / (ms_e2e * 1e-3), 'unit': 'images/s', 'h2d_bytes_per_step': B_PER_GPU * 3 * IMG * IMG * 4,
                    'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': launches,
            'clocks': sampler.summary(),
            'roofline': roof,
            'kernel_breakdown': breakdown,
            'cpu_baseline': cpu,
        }
        print(json.dumps(line))
    if world > 1:
        # Tear-down: drop the captured graph (it holds the NCCL communicator's captured work) before leaving, and do not
        # call destroy_process_group() -- with a live captured all-reduce it can block for the watchdog timeout.
        barrier()
        ts.graph = None
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='engine', choices=['engine', 'reference'])
    ap.add_argument('--no-cpu-baseline', dest='no_cpu_baseline', action='store_true')
    ap.add_argument('--no-graph', dest='no_graph', action='store_true')
    ap.add_argument('--no-overlap', dest='no_overlap', action='store_true',
                    help='one gradient all-reduce after the backward pass instead of bucketed collectives under it (A/B)')
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS))
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_engine(args)


if __name__ == '__main__':
    main()
