#!/usr/bin/env python
"""Benchmark of the fusion-FPN training step (forward + loss + backward + SGD) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Metric = BASELINE.json's "fwd+bwd samples/sec"; workload at N=1 = configs[1]
(C2: FPNHybridFusion, batch 8, image 1x32x128x128, SLO 320x128, crop relative_2d_max, bf16 storage / fp32
accumulate), weak-scaled (8 samples per GPU) for N>1, one process per GPU over NCCL.
`value`  : steps timed with inputs resident in HBM (CUDA events, barrier + synchronize on both sides, max
           over ranks).  The per-step working set (several GB of activations) is far larger than the 126 MB
           L2, so no explicit L2 flush is needed between iterations.
`e2e`    : same step driven from pinned HOST buffers through the public module API: H2D copy of image / slo
           / mask every step and a D2H read of the loss inside the timed region.
`roofline`: the dominant kernel timed alone with CUDA events, algorithmic bytes / time vs MEASURED_PEAKS.json.
`cpu_baseline` / `--impl reference`: the oracle port (torch CPU ops restating the reference) on the host cores.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, 'multimodal-fusion-fpn_b200')       # put on sys.path by the GPU arm only: the reference arm must import
                                                             # oracle/_ref's `models` / `config`, never this repository's

# BASELINE.json configs.  model_elems = conv-boundary traffic model, elements read + written per sample, forward.
WORKLOADS = {
    'C1': dict(name='C1: FPNHybridFusion, batch 1, image 1x64x128x128, slo 128x128, crop relative_2d_max', B=1, S=64, H=128, W=128,
               S2=128, W2=128, model_elems=664.8e6, scaling='weak'),
    'C2': dict(name='C2: FPNHybridFusion GA segmentation, batch 8/GPU, image 1x32x128x128, slo 320x128, crop relative_2d_max',
               B=8, S=32, H=128, W=128, S2=320, W2=128, model_elems=367.2e6, scaling='weak'),
    'C3': dict(name='C3: Level5 fusion FPN at full B-scan depth, batch 8/GPU, image 1x32x496x128, slo 320x128, crop relative_2d_max',
               B=8, S=32, H=496, W=128, S2=320, W2=128, model_elems=1296.8e6, scaling='weak'),
    'C4': dict(name='C4: vessel segmentation shapes, batch 16/GPU, image 1x32x128x128, slo 320x128, crop relative_2d_max',
               B=16, S=32, H=128, W=128, S2=320, W2=128, model_elems=367.2e6, scaling='weak'),
    'C5': dict(name='C5: data-parallel sweep, GLOBAL batch 64 split over the GPUs, image 1x32x128x128, slo 320x128, '
                    'crop relative_2d_max', B=64, S=32, H=128, W=128, S2=320, W2=128, model_elems=367.2e6, scaling='strong'),
}
WORKLOAD = dict(WORKLOADS['C2'])
METRIC, UNIT = 'fwd+bwd samples/sec', 'samples/s'


def select_workload(name, world):
    w = dict(WORKLOADS[name])
    w['id'] = name
    if name == 'C5':
        assert 64 % world == 0, 'C5 splits a global batch of 64'
        w['B'] = 64 // world
    WORKLOAD.clear()
    WORKLOAD.update(w)
    return WORKLOAD


def cpu_model():
    try:
        with open('/proc/cpuinfo') as f:
            for line in f:
                if line.startswith('model name'):
                    return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get('hbm_gbs', 6650.0), 'measured (MEASURED_PEAKS.json hbm_gbs, burst copy)'
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled from a
    thread every 10 ms (nvidia-smi -lms takes longer to start than a short timed region lasts); falls back to one
    nvidia-smi query if NVML is unavailable."""
    REASONS = (('hw_slowdown', 0x8), ('sw_thermal_slowdown', 0x20), ('hw_thermal_slowdown', 0x40), ('sw_power_cap', 0x4))

    def __init__(self, gpu_index):
        self.idx, self.samples, self.reasons, self.max_mhz = gpu_index, [], set(), None
        self._stop, self._thread, self._nvml = threading.Event(), None, None

    def _poll(self):
        nv, h = self._nvml
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = int(visible.split(',')[self.idx]) if visible and visible.split(',')[self.idx].isdigit() else self.idx
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self._nvml = (nv, h)
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._nvml = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self._nvml is not None:
            self._stop.set()
            self._thread.join(timeout=2)
            if self.samples:
                sm = sorted(self.samples)
                out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(sm),
                           source='nvml, 10 ms poll during the timed region')
                return out
        try:
            q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
                 'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
            r = subprocess.run(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-i', str(self.idx)],
                               capture_output=True, text=True, timeout=10).stdout.strip().split(',')
            out.update(sm_mhz=float(r[0]), sm_max_mhz=float(r[1]), samples=1, source='nvidia-smi, one query after the timed region',
                       reasons=[n for n, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6])
                                if v.strip().lower().startswith('active')])
        except Exception:
            out['reasons'] = ['clock query unavailable']
        return out


def _host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


REF_DIR = os.path.join(ROOT, 'oracle', '_ref')


def _reference_step_fn(batch_size):
    """-> (kind, step()).  'reference' = the UNMODIFIED reference modules copied to oracle/_ref by oracle/make_ref.py
    (FPNHybridFusion + weight_init + Mix loss, SURVEY.md section 8c import recipe); 'port' = the oracle restatement, only when
    that copy is missing.  One step = forward + loss + backward of ``batch_size`` samples of the workload shape, fp32, all host
    threads.  Nothing of this repository's product is imported here."""
    import torch
    threads = _host_threads()
    torch.set_num_threads(threads)
    w = WORKLOAD
    from oracle import fusion_fpn_oracle as O            # synthetic batch generator (and the fallback port)
    batch = O.synthetic_batch(batch_size, w['S'], w['H'], w['W'], w['S2'], w['W2'], seed=1234)
    if os.path.isdir(os.path.join(REF_DIR, 'models')):
        saved_cwd, saved_argv = os.getcwd(), sys.argv
        os.chdir(REF_DIR)                                  # the .ini is read cwd-relative (fusion_nets.py:24-26)
        sys.path[:] = [q for q in sys.path if os.path.abspath(q or '.') != PKG]
        sys.path.insert(0, REF_DIR)
        sys.argv = ['x', '--training-dataset', 'hrf_fusion', '--model', 'FPNHybridFusion', '--fusion-modality', 'slo',
                    '--crop', 'relative_2d_max']
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                import config as _cfg                      # noqa: F401
                from models.fusion_nets import factory_classes
                from common import loss as rloss, weight_init as rinit
                torch.manual_seed(1234)                    # train.py:42
                model = factory_classes['FPNHybridFusion']()
                model.apply(rinit.weight_init)             # train.py:56
            import models.fusion_nets as _fn
            assert os.path.abspath(_fn.__file__).startswith(REF_DIR), 'the reference arm must run the reference modules'
        finally:
            sys.argv = saved_argv
            os.chdir(saved_cwd)
        model.train()
        crit = rloss.Mix({'Dice': rloss.Dice_loss_jointv2('prediction', 'mask'), 'BCE': rloss.BCE_Lossv2('prediction', 'mask')})

        def step():
            model.zero_grad(set_to_none=True)
            loss, _ = crit(batch, model(batch))
            loss.backward()
            return float(loss)
        return 'reference', step, threads
    sd = O.make_state_dict(seed=1234)

    def step():
        return float(O.loss_and_grads(sd, batch)[0])
    return 'port', step, threads


def _reference_batch():
    """The GPU arm's per-step batch when the host has the memory for it (stock PyTorch keeps ~2.6 GB of activations per C2
    sample), else the largest power-of-two batch that fits."""
    w = WORKLOAD
    per_sample_gb = 2 * 4 * w['model_elems'] * 2.0 / 1e9 / 2          # ~2x the traffic model's elements stay alive, fp32
    try:
        import psutil
        avail = psutil.virtual_memory().available / 1e9
    except Exception:
        avail = 32.0
    b = w['B']
    while b > 1 and b * per_sample_gb > 0.5 * avail:
        b //= 2
    return b


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = WORKLOAD
    bsz = args.ref_batch or _reference_batch()
    steps, warmup = args.steps, args.warmup
    # bounded: the whole run must end within a few minutes on the host cores -> probe one step, then cut the step count if needed
    kind, step, threads = _reference_step_fn(bsz)
    t0 = time.perf_counter()
    step()
    probe = time.perf_counter() - t0
    budget = 240.0
    if probe * (steps + warmup) > budget:
        warmup = min(warmup, 1)
        steps = max(2, min(steps, int(budget / probe) - warmup))
    ts = []
    for i in range(max(warmup - 1, 0) + steps):                        # the probe was the first warm-up step
        t0 = time.perf_counter()
        step()
        if i >= max(warmup - 1, 0):
            ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    rate = bsz / sec
    import torch
    what = ('the unmodified reference (oracle/_ref: models/fusion_nets.py FPNHybridFusion + common/loss.py Mix, weight_init, seed 1234)'
            if kind == 'reference' else 'oracle port of the reference (oracle/_ref missing)')
    sample = (f'{what}, torch {torch.__version__} CPU fp32, {threads} host threads on {cpu_model()}; each step = fwd+loss+bwd of '
              f'{bsz} sample(s) of the {w["id"]} shape; {steps} timed after {warmup} warm-up')
    line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': world, 'steps': steps,
            'warmup': warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': w['scaling'], 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': w['name'], 'per_gpu_batch': w['B'], 'per_step_samples': bsz, 'cpu_model': cpu_model()},
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def kernel_roofline(torch, ops, peak_gbs, peak_src, iters=20):
    """Time the hot kernels alone at the C2 shapes (tools/kernel_cases.py: the same named cases the ncu captures under profiles/
    use): CUDA events on the launching stream, inputs >> L2.  Convolutions run with their packed weights in the arena, as in the
    training step (the timed region holds the kernel, not a per-call weight-packing launch the step does not make either)."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import kernel_cases as KC
    cases = KC.build(torch, ops)
    tflops_peak = 1653.8
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        with open(pk) as f:
            tflops_peak = json.load(f).get('bf16_tflops', tflops_peak)             # burst figure: kernels timed alone
    traffic = {}
    tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f)
    arena_buf = torch.empty((16 << 20) + 1024, dtype=torch.uint8, device='cuda')
    arena = arena_buf[(-arena_buf.data_ptr()) % 1024:][:16 << 20]          # the library wants a 1 KiB-aligned buffer
    dev_index = torch.cuda.current_device()
    res = []
    for name in KC.DEFAULT_BENCH:
        fn, nbytes, flops = cases[name]['make']()
        ops.weight_arena_begin(arena)                     # record this case's weight image, then replay from the arena
        fn()
        ops.weight_arena_seal(dev_index)
        ops.weight_arena_enable(dev_index, True)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ops.weight_arena_end(dev_index)
        sec = e0.elapsed_time(e1) / iters * 1e-3
        ai = flops / max(nbytes, 1)
        bound = 'tensor' if ai > KC.RIDGE_FLOP_PER_BYTE else 'hbm'
        gbs, tfs = nbytes / sec / 1e9, flops / sec / 1e12
        t = traffic.get(name, {})
        res.append(dict(kernel=name, note=cases[name]['note'], bound=bound, sec=sec, bytes=nbytes, flops=flops, flop_per_byte=ai,
                        gbs=gbs, tflops=tfs, frac=(tfs / tflops_peak) if bound == 'tensor' else (gbs / peak_gbs),
                        frac_hbm=gbs / peak_gbs, frac_tensor=tfs / tflops_peak,
                        ncu=({k: t[k] for k in ('dram_bytes', 'duration_us', 'tensor_pipe_pct', 'dram_pct', 'file') if k in t} or None)))
        del fn
        torch.cuda.empty_cache()
    return res, tflops_peak


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='C2', choices=sorted(WORKLOADS), help='BASELINE.json workload (default C2 = the metric\'s)')
    ap.add_argument('--dtype', default='bf16', choices=['bf16', 'f32'])
    ap.add_argument('--ref-batch', type=int, default=0, help='reference arm: samples per step (default: the GPU arm\'s batch)')
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-kernel-roofline', action='store_true')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    select_workload(args.config, world)
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), 'bench.py needs a B200: there is no CPU path for the product'
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    sys.path.insert(0, PKG)
    from __graft_entry__ import import_mirror
    cfg, fusion_nets, loss_mod, weight_init = import_mirror()
    import ffpn
    from ffpn import ops
    from ffpn.trainer import FusionTrainer
    from oracle import fusion_fpn_oracle as O            # synthetic batch generator + cpu_baseline only

    dtype = torch.bfloat16 if args.dtype == 'bf16' else torch.float32
    ffpn.set_compute_dtype(dtype)
    torch.manual_seed(1234)                                # train.py:42
    with contextlib.redirect_stdout(io.StringIO()):
        model = fusion_nets.factory_classes['FPNHybridFusion']()
    model.apply(weight_init.weight_init)                   # train.py:56
    model = model.cuda().train()
    crit = loss_mod.Mix({'Dice': loss_mod.Dice_loss_jointv2('prediction', 'mask'),
                         'BCE': loss_mod.BCE_Lossv2('prediction', 'mask')})
    w = WORKLOAD
    host = O.synthetic_batch(w['B'], w['S'], w['H'], w['W'], w['S2'], w['W2'], seed=1234 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    dev = {k: v.cuda(non_blocking=True) for k, v in host.items()}
    trainer = FusionTrainer(model, crit, lr=cfg.learning_rate, momentum=0.9, weight_decay=1e-4)

    if ffpn.lib.is_debug_build():
        raise SystemExit('bench.py refuses the -DFFPN_DEBUG library (FFPN_LIB=debug): its ablation switches invalidate results')
    trainer.step(dev)                                      # first eager step: records the packed-weight arena
    torch.cuda.synchronize()
    n0, r0 = ffpn.lib.launch_count(local_rank), ffpn.lib.route_counts(local_rank)
    trainer.step(dev)                                      # steady-state eager step: counts our launches per step
    torch.cuda.synchronize()
    launches_per_step = ffpn.lib.launch_count(local_rank) - n0
    routes = {k: v - r0[k] for k, v in ffpn.lib.route_counts(local_rank).items()}      # conv calls per kernel family, one step
    if args.dtype == 'bf16' and (routes['cuda_core'] != 0 or routes['tcgen05_gen1'] != 0):
        raise SystemExit(f'bench.py: bf16 conv calls left the warp-specialised tcgen05 / stem kernels: {routes}')
    use_graph = not args.no_graph
    if use_graph:
        trainer.capture(dev)
    step = (lambda b=None: trainer.replay(b)) if use_graph else (lambda b=None: trainer.step(b if b is not None else dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    # ---- end to end: pinned host -> device every step, loss read back every step ---------------------------
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    if use_graph:                                          # untimed: create the copy stream / staging buffers of the prefetch path
        trainer.prefetch(host)
        trainer.replay(prefetched=True)
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = 0.0
    if use_graph:
        # every step's batch goes pinned host -> device inside the timed region; the copy of step i+1 is issued on a copy
        # stream while step i runs (FusionTrainer.prefetch), and every step's loss is read back
        trainer.prefetch(host)
        for i in range(args.steps):
            loss = trainer.replay(prefetched=True)
            if i + 1 < args.steps:
                trainer.prefetch(host)
            last = float(loss.item())
    else:
        for _ in range(args.steps):
            last = float(step({k: v.cuda(non_blocking=True) for k, v in host.items()}).item())
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t = torch.tensor([ms, ms_e2e], device='cuda', dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    samples = w['B'] * world * args.steps
    value, e2e_value = samples / (ms * 1e-3), samples / (ms_e2e * 1e-3)

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': w['scaling'], 'vs_baseline': None,
            'dtype': args.dtype, 'data': 'synthetic',
            'config': {'workload': w['name'], 'per_gpu_batch': w['B'], 'global_batch': w['B'] * world,
                       'parallelism': f'dp{world}', 'cuda_graph': use_graph,
                       'branch_streams': os.environ.get('FFPN_STREAMS', '1'), 'pdl': os.environ.get('FFPN_PDL', '1'),
                       'optimizer': 'SGD(0.1, 0.9, wd 1e-4) fused',
                       'l2': 'per-step working set >> 126 MB L2 (no flush needed)', 'final_loss': last},
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'ms_per_step': ms_e2e / args.steps},
            'gpu_launches': int(launches_per_step * args.steps), 'conv_routes_per_step': routes, 'build': 'release'}
    if rank == 0:
        peak, src = peaks()
        if not args.no_kernel_roofline:
            trainer.close()                                 # detach the trainer's arena: the kernel timings below use their own
            ks, tflops_peak = kernel_roofline(torch, ops, peak, src)
            top = ks[0]                                     # conv_fwd_l1: the kernel with the largest share of the step
            line['roofline'] = {'bound': top['bound'], 'kernel': top['kernel'] + ': ' + top['note'], 'achieved': top['gbs'], 'peak': peak,
                                'unit': 'GB/s', 'frac': top['frac'], 'traffic': (top['ncu'] or {}).get('dram_bytes'),
                                'traffic_source': (top['ncu'] or {}).get('file'), 'peak_source': src,
                                'algorithmic_bytes_per_launch': top['bytes'], 'sec_per_launch': top['sec']}
            line['kernels'] = [{k: r[k] for k in ('kernel', 'bound', 'gbs', 'tflops', 'frac', 'frac_hbm', 'frac_tensor', 'sec', 'bytes',
                                                  'flops', 'flop_per_byte', 'ncu')} for r in ks]
            line['peaks'] = {'hbm_gbs': peak, 'bf16_tflops_burst': tflops_peak, 'ridge_flop_per_byte': 212.0}
            # whole-step roofline: conv-boundary traffic model of SURVEY.md section 8d (bf16, fwd+bwd = 3 x fwd)
            step_bytes = w['model_elems'] * 2 * 3 * w['B']
            line['step_roofline'] = {'model_bytes_per_step': step_bytes, 'achieved_gbs': step_bytes / (ms / args.steps * 1e-3) / 1e9,
                                     'frac': step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak}
        if world == 1 and not args.no_cpu_baseline:
            # the reference's modules share their import names (config, models, common) with this repository's mirror, so the
            # bounded CPU sample runs in a fresh interpreter: bench.py --impl reference at batch 1, 3 timed steps after 1 warm-up
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--config', w['id'], '--steps', '3',
                                '--warmup', '1', '--ref-batch', '1'], capture_output=True, text=True, timeout=900,
                               env={k: v for k, v in os.environ.items() if k not in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK')})
            try:
                ref_line = json.loads(r.stdout.strip().splitlines()[-1])
                line['cpu_baseline'] = dict(ref_line['cpu_baseline'], cpu_model=cpu_model())
            except Exception:
                line['cpu_baseline'] = {'value': None, 'unit': UNIT, 'cores': _host_threads(), 'kind': 'reference',
                                        'sample': 'failed: ' + (r.stderr.strip().splitlines() or ['no output'])[-1][:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        # The captured step holds NCCL kernels; tearing the communicator down under a live CUDA graph can block at exit.  Every
        # rank has printed / reduced what it had to: synchronise, then leave without the destructor chain.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
