#!/usr/bin/env python
"""Headline benchmark: MPix/s of encoder + context-model (probclass) forward.

    python bench.py --gpus N --steps K --warmup W [--workload kodak24|b64_512|cfg1|cfg4] [--mode fp32|exact|fast]
    python bench.py --impl reference ...      # the restated CPU path on the host cores

One "step" = ae.encode(x) + pc.bitcost(qbar, symbols) over one batch of synthetic
uint8 images (BASELINE.json metric "MPix/s encode+probclass fwd").  Prints ONE
JSON line (see DESIGN.md "Measurement").  N>1: launched under torchrun, one rank
per GPU, images sharded by batch (weak scaling: every rank runs the full
per-GPU batch), one NCCL all-reduce of the metric sums per step.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (ae config, N per GPU, H, W)      BASELINE.json configs
    'cfg1': ('cvpr/low', 1, 128, 128),         # configs[0]
    'kodak24': ('cvpr/low', 24, 768, 512),     # configs[1]  (default: the config the metric is quoted on)
    'b64_512': ('cvpr/low', 64, 512, 512),     # north_star headline shape
    'cfg4': ('cvpr/hi', 32, 512, 512),         # configs[3], per-GPU share of B=256
}
L2_FLUSH_BYTES = 256 << 20


def encoder_macs_per_pixel(C):
    """h1 1200 + h2 12800 + 32 x 9216 + to_bn 25*128*(C+1)/64."""
    return 1200 + 12800 + 32 * 9216 + 25 * 128 * (C + 1) / 64.0


def probclass_macs(C, h, w, k=24, L=6):
    """non-masked taps, exact per-layer output shapes incl. halo."""
    d0 = (C + 3) * (h + 6) * (w + 6) * 13 * k
    d1 = (C + 2) * (h + 4) * (w + 4) * 14 * k * k
    d2 = (C + 1) * (h + 2) * (w + 2) * 14 * k * k
    d3 = C * h * w * 14 * k * L
    return d0 + d1 + d2 + d3


def load_traffic(workload, mode):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture of the same workload / mode (profiles/r2_conv3x3_traffic.json, written by tools/gpu_profile.sh from the
    .ncu-rep of this round's kernel); None for other workloads / modes."""
    p = os.path.join(ROOT, 'profiles', 'r2_conv3x3_traffic.json')
    if not os.path.exists(p):
        p = os.path.join(ROOT, 'profiles', 'r1b_conv3x3_traffic.json')
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        d = json.load(f)
    if d.get('workload') != workload or d.get('mode') != mode:
        return None, None
    return d['traffic_bytes_per_launch'], d['source']


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.Q,
                                       '--format=csv,noheader,nounits', '-lms', '100'],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ''
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


def make_models(ae_name, mode):
    from imgcomp_cvpr_b200 import autoencoder, config, probclass, weights
    a, p = config.ae_config(ae_name), config.pc_config('cvpr/res_shallow')
    W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    ae = autoencoder.get_network_cls(a)(a, weights=W, mode=mode)
    pc = probclass.get_network_cls(p)(p, num_centers=a.num_centers, weights=W)
    return a, p, W, ae, pc


def cpu_reference_step(x_u8, W, C):
    """The restated CPU path (oracle, torch-CPU conv kernels): encode + probclass bitcost."""
    from oracle import imgcomp_oracle as O
    enc = O.encode(x_u8.astype(np.float32), W, C)
    bc, _ = O.pc_bitcost(enc['qbar'], enc['symbols'], W, W['autoencoder/encoder/centers'][0])
    return float(bc.sum())


def pick_cpu_threads(xs, W, C):
    """torch-CPU convs do not scale to every core of a many-core host: time one pass at a few
    thread counts and keep the fastest (reported as `cores`)."""
    import torch
    best, best_t = None, None
    for n in sorted({os.cpu_count(), min(os.cpu_count(), 32), min(os.cpu_count(), 16)}, reverse=True):
        torch.set_num_threads(n)
        cpu_reference_step(xs, W, C)              # warm-up at this thread count
        t0 = time.perf_counter()
        cpu_reference_step(xs, W, C)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def run_reference(args):
    """--impl reference: the reference's CPU path is TF-1.4 and cannot run here;
    the oracle's restatement is timed on the host cores on a bounded sample of the workload."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import torch
    from imgcomp_cvpr_b200 import config, weights
    from oracle import imgcomp_oracle as O
    ae_name, N, H, Wd = WORKLOADS[args.workload]
    a, p = config.ae_config(ae_name), config.pc_config('cvpr/res_shallow')
    W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    O.set_backend('torch')
    n_sample = 1
    x = weights.synthetic_images(n_sample, H, Wd, seed=1234)
    cores = pick_cpu_threads(x, W, a.num_chan_bn)
    for _ in range(args.warmup):
        cpu_reference_step(x, W, a.num_chan_bn)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(x, W, a.num_chan_bn)
    dt = (time.perf_counter() - t0) / args.steps
    val = n_sample * H * Wd / dt / 1e6
    sample = '%d of %d images of %dx%d per step' % (n_sample, N, H, Wd)
    print(json.dumps({
        'impl': 'reference', 'metric': 'MPix/s encode+probclass fwd', 'value': val, 'unit': 'MPix/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'ae': ae_name, 'pc': 'cvpr/res_shallow', 'batch_per_gpu': N,
                   'H': H, 'W': Wd, 'sample': sample},
        'cpu_baseline': {'value': val, 'unit': 'MPix/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': val, 'unit': 'MPix/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }))


def oracle_parity(ae, pc, ae_name, mode, x_u8_dev, W, C, nchk=1):
    """Parity evidence on THIS workload against the CPU oracle (oracle/imgcomp_oracle.py, torch-CPU backend): the first
    `nchk` images through code/val.py:81-89 -- symbols, bpp -- for the timed mode AND the library's float32 FFMA mode.
    `safe` = positions whose float64 latent is further than eps from a quantizer decision boundary (there a symbol may
    not differ between two float32-class implementations)."""
    import torch
    from oracle import imgcomp_oracle as O
    xs = x_u8_dev[:nchk].contiguous()
    x_np = xs.cpu().numpy()
    H, Wd = x_np.shape[2], x_np.shape[3]
    centers = W['autoencoder/encoder/centers']
    O.set_backend('torch')
    try:
        t0 = time.perf_counter()
        enc = O.encode(x_np.astype(np.float32), W, C)
        bc, _ = O.pc_bitcost(enc['qbar'], enc['symbols'], W, centers[0])
        z64 = O.encode(x_np.astype(np.float64), W, C, dtype=np.float64)['z']
        oracle_s = time.perf_counter() - t0
    finally:
        O.set_backend('numpy')
    o_bpp = bc.reshape(nchk, -1).sum(axis=1, dtype=np.float64) / (H * Wd)
    c = np.sort(centers.astype(np.float64))
    margin = np.abs(z64[..., None] - (c[1:] + c[:-1]) / 2).min(axis=-1)
    sym64 = np.abs(z64[..., None] - centers.astype(np.float64)).argmin(axis=-1)
    out = {'against': 'oracle/imgcomp_oracle.py (float32 restatement of code/val.py:81-89, torch-CPU convs) + its float64 latent',
           'images': nchk, 'symbols': int(enc['symbols'].size), 'oracle_seconds': oracle_s,
           'oracle_f32_vs_f64_symbol_mismatches': int((enc['symbols'] != sym64).sum()), 'bpp_oracle': o_bpp.tolist()}
    eps = {'fp32': 1e-4, 'exact': 2e-4, 'fast': 2e-2}
    modes = [mode] + ([] if mode == 'fp32' else ['fp32'])
    for m in modes:
        if m == mode:
            a_m, p_m = ae, pc
        else:
            _, _, _, a_m, p_m = make_models(ae_name, m)
        e = a_m.encode(xs, is_training=False)
        p_m.bitcost(e.qbar, e.symbols, is_training=False, pad_value=p_m.auto_pad_value(a_m))
        bpp = (p_m.last_bits_per_image / (H * Wd)).cpu().numpy()
        sym = e.symbols.cpu().numpy()
        z = e.z.cpu().numpy()
        mism = sym != enc['symbols']
        safe = margin > eps[m]
        out[m] = {'symbol_mismatches': int(mism.sum()), 'symbol_mismatches_safe_set': int((mism & safe).sum()),
                  'safe_eps': eps[m], 'symbol_mismatches_vs_float64': int((sym != sym64).sum()),
                  'max_abs_dz_vs_float64': float(np.abs(z - z64).max()), 'mean_abs_dz_vs_float64': float(np.abs(z - z64).mean()),
                  'max_abs_dbpp': float(np.abs(bpp - o_bpp).max()), 'bpp': bpp.tolist()}
        del e
        if m != mode:
            del a_m, p_m
            torch.cuda.empty_cache()
    out['symbol_mismatches'] = out[mode]['symbol_mismatches']
    out['max_abs_dbpp'] = out[mode]['max_abs_dbpp']
    return out


def measure_encode_pc(workload, mode, steps, warmup, world, rank, local, with_parity, sample_clocks=True):
    """One workload of the headline metric: K timed steps of ae.encode + pc.bitcost (device-resident `value`), the same
    K steps with the library's per-launch events (roofline), and the end-to-end pass from pinned host memory."""
    import torch
    import torch.distributed as dist
    from imgcomp_cvpr_b200 import _lib, weights
    L = _lib.lib()
    ae_name, N, H, Wd = WORKLOADS[workload]
    a, p, W, ae, pc = make_models(ae_name, mode)
    C = a.num_chan_bn
    # every rank gets its own shard of the (virtual) global batch: different seeds per rank
    x_host = torch.from_numpy(weights.synthetic_images(N, H, Wd, seed=1234 + rank)).pin_memory()
    x_dev = x_host.cuda()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device='cuda')
    metric = torch.zeros(2, dtype=torch.float64, device='cuda')
    bits_host = torch.empty(N, dtype=torch.float64).pin_memory()
    sym_host = torch.empty((N, C, H // 8, Wd // 8), dtype=torch.uint8).pin_memory()
    last = {}

    def step(x):
        enc = ae.encode(x, is_training=False)
        last['sym8'] = ae.extra['symbols_u8']
        pc.bitcost(enc.qbar, enc.symbols, is_training=False, pad_value=pc.auto_pad_value(ae))
        if world > 1:      # the final metric all-reduce of a sharded val run: [sum bits, sum pixels]
            metric[0] = pc.last_bits_per_image.sum()
            metric[1] = float(N * H * Wd)
            dist.all_reduce(metric)
        return pc.last_bits_per_image

    def timed(fn, k):
        """K steps, each bracketed by its own CUDA-event pair on the launching stream; the L2 flush between
        steps is enqueued outside the pairs (not timed); nothing synchronises with the host until the end."""
        evs = []
        for _ in range(k):
            flush.fill_(1)                     # L2 flush between timed iterations
            a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_ev.record()
            fn()
            b_ev.record()
            evs.append((a_ev, b_ev))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / k

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step(x_dev)
    barrier()
    parity = None
    if with_parity and rank == 0:
        parity = oracle_parity(ae, pc, ae_name, mode, x_dev, W, C)
    barrier()
    # EVERY rank (step() holds the all-reduce): rank 0's emptied allocator cache makes its next step cudaMalloc the workspace
    # again, which must not land in the timed pass
    for _ in range(2):
        step(x_dev)
    barrier()
    # ---- device-resident throughput (`value`): K steps, nothing else on the stream
    launches0 = L.ic_launch_count()
    sampler = ClockSampler(local) if (rank == 0 and sample_clocks) else None
    ms = timed(lambda: step(x_dev), steps)
    barrier()
    clocks = sampler.stop() if sampler else None
    launches = L.ic_launch_count() - launches0
    # ---- live per-kernel-class timing (roofline): a second pass of K steps (all ranks) with every launch of the library
    # bracketed by CUDA events on the launching stream; kept out of the pass above so that the timed pass carries
    # nothing but the path.
    L.ic_profile_reset()
    L.ic_profile_enable(1)
    ms_profiled = timed(lambda: step(x_dev), steps)
    barrier()
    L.ic_profile_enable(0)
    prof = {}
    for cls, name in enumerate(['conv3x3', 'conv_other', 'elementwise', 'probclass', 'msssim']):
        t, n = _lib.c_double(), _lib.c_longlong()
        _lib.check(L.ic_profile_get(cls, t, n))
        prof[name] = (t.value, n.value)
    # ---- end to end: pinned host uint8 -> H2D -> step -> D2H of the per-image bit sums AND the symbols (uint8: what a
    # codec-style consumer of the path needs on the host)
    # Double-buffered like a streaming val driver: the H2D copy of batch i+1 (copy stream) overlaps the compute of
    # batch i; every step still copies its own inputs from pinned host memory and reads its result back, all inside
    # the timed region.  (No explicit L2 flush here: each step touches > 2 GB of activations, far beyond the 126 MB L2.)
    copy_stream = torch.cuda.Stream()
    x_host2 = x_host.clone().pin_memory()
    host_bufs = [x_host, x_host2]
    dev_bufs = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    copied = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def issue_copy(i):
        b = i & 1
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[b])          # the buffer's previous batch has been encoded
            dev_bufs[b].copy_(host_bufs[b], non_blocking=True)
            copied[b].record(copy_stream)

    def e2e_run(k):
        main_s = torch.cuda.current_stream()
        for b in range(2):
            consumed[b].record(main_s)
        issue_copy(0)
        for i in range(k):
            b = i & 1
            if i + 1 < k:
                issue_copy(i + 1)
            main_s.wait_event(copied[b])
            bits = step(dev_bufs[b])
            consumed[b].record(main_s)
            bits_host.copy_(bits, non_blocking=True)
            sym_host.copy_(last['sym8'], non_blocking=True)
    e2e_run(2)
    barrier()
    a_ev, b_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a_ev.record()
    e2e_run(steps)
    b_ev.record()
    b_ev.synchronize()
    ms_e2e = a_ev.elapsed_time(b_ev) / steps
    barrier()
    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = t.tolist()
    pix = N * H * Wd
    peaks, peak_src = load_peaks()
    # roofline of the dominant kernel class: the 32 3x3 128->128 convs
    t3, n3 = prof['conv3x3']
    m_rows = N * (H // 4) * (Wd // 4)
    flop_per_launch = 2.0 * m_rows * 128 * 1152
    avg_ms = t3 / max(n3, 1)
    achieved = flop_per_launch / (avg_ms * 1e-3) / 1e12 if n3 else 0.0
    peak = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops'))
    traffic, traffic_src = load_traffic(workload, mode)
    step_flop = 2.0 * (encoder_macs_per_pixel(C) * pix + N * probclass_macs(C, H // 8, Wd // 8, p.arch_param__k))
    mmas = {'exact': 3, 'fast': 1, 'fp32': 0}[mode]
    res = {
        'value': world * pix / (ms * 1e-3) / 1e6, 'ms_per_step': ms,
        'config': {'workload': workload, 'ae': ae_name, 'pc': 'cvpr/res_shallow', 'batch_per_gpu': N,
                   'H': H, 'W': Wd, 'mode': mode, 'parallelism': 'batch-shard x%d' % world,
                   'l2': 'flushed between timed iterations (%d MiB write)' % (L2_FLUSH_BYTES >> 20)},
        'e2e': {'value': world * pix / (ms_e2e * 1e-3) / 1e6, 'unit': 'MPix/s', 'h2d_bytes_per_step': int(x_host.numel()),
                'd2h_bytes_per_step': int(bits_host.numel() * 8 + sym_host.numel()), 'ms_per_step': ms_e2e,
                'd2h': 'per-image bit sums (float64) + symbols (uint8)'},
        'gpu_launches': int(launches), 'clocks': clocks,
        'roofline': {'bound': 'tensor', 'kernel': 'conv3x3_128x128 (%s)' % mode, 'achieved': achieved,
                     'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak if peak else None,
                     'peak_source': '%s bf16_tflops_sustained' % peak_src, 'traffic': traffic,
                     'traffic_unit': 'bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)', 'traffic_source': traffic_src,
                     # in + out activations (hi/lo fp16 planes = 4 B / element) + 0 / 1 / 2 residual reads, averaged over a residual group
                     'algorithmic_bytes_per_launch': m_rows * 128 * 4.0 * (2 * 6 + 3 + 1) / 6.0,
                     'flop_per_launch': flop_per_launch, 'avg_launch_ms': avg_ms, 'launches': n3,
                     'share_of_step': t3 / (ms_profiled * steps) if ms_profiled else None,
                     'ms_per_step_with_launch_events': ms_profiled,
                     'mma_flops_per_algorithmic_flop': mmas,
                     'tensor_issue_frac': (achieved * mmas / peak) if peak else None,
                     'note': 'exact = fp16x3 split: 3 tensor-core FLOPs per algorithmic FLOP, so frac <= 1/3 by construction',
                     'step_algorithmic_tflop': step_flop / 1e12,
                     'step_tflops': step_flop / (ms * 1e-3) / 1e12,
                     'step_frac': step_flop / (ms * 1e-3) / 1e12 / peak if peak else None},
        'kernel_ms_per_step': {k: v[0] / steps for k, v in prof.items()},
        'parity': parity,
    }
    ctx = {'ae': ae, 'pc': pc, 'W': W, 'a': a, 'p': p, 'x_dev': x_dev, 'timed': timed, 'barrier': barrier}
    return res, ctx


def measure_train_step(steps):
    """BASELINE.json configs[2]: one training step (forward + MS-SSIM / rate loss + backward + two Adam groups) at B = 32,
    160x160, cvpr/med + res_shallow, replayed as one CUDA graph (code/train.py:86-132,252,339-349)."""
    import torch
    from imgcomp_cvpr_b200 import _lib, config, trainer, weights
    a, p = config.ae_config('cvpr/med'), config.pc_config('cvpr/res_shallow')
    W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    B, S = 32, 160
    x = torch.from_numpy(weights.synthetic_images(B, S, S, seed=77)).cuda()
    tr = trainer.Trainer(a, p, W, num_itr_per_epoch=1000, mode='exact')
    L = _lib.lib()
    n0 = L.ic_launch_count()
    tr.step(x)                                    # eager once: counts the launches one step is made of
    per_step = L.ic_launch_count() - n0
    tr.enable_cuda_graph(x)
    for _ in range(3):
        out = tr.step(x)
    torch.cuda.synchronize()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device='cuda')
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = tr.step(x)
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a_.elapsed_time(b_) for a_, b_ in evs) / steps
    pix = B * S * S
    # forward MACs of SURVEY.md 8(d) (encoder + decoder + context model) x 3: forward, data gradient, filter gradient
    tflop = 3 * 2 * (310562 + 309488 + 10470) * pix / 1e12
    peaks, peak_src = load_peaks()
    peak = peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops'))
    del tr
    torch.cuda.empty_cache()
    return {'workload': 'cfg3: B=32 160x160 cvpr/med + res_shallow, forward + loss + backward + Adam (one CUDA graph replay)',
            'ms_per_step': ms, 'images_per_s': B / (ms * 1e-3), 'MPix_per_s': pix / (ms * 1e-3) / 1e6,
            'dtype': 'f32 + f16x3 tcgen05 convs (3x3 trunk, h2, h12, h13, context-model layers 1-3: forward, data gradient, filter gradient)', 'steps': steps,
            'gpu_launches_per_step': int(per_step), 'loss': out['total_loss'], 'bpp': out['bpp'], 'ms_ssim': out['ms_ssim'],
            'roofline': {'bound': 'tensor', 'achieved': tflop / (ms * 1e-3), 'peak': peak, 'unit': 'TFLOP/s',
                         'frac': tflop / (ms * 1e-3) / peak if peak else None, 'peak_source': '%s bf16_tflops_sustained' % peak_src,
                         'algorithmic_tflop_per_step': tflop}}


def measure_real_bpp(n_images=8):
    """BASELINE.json configs[4] (--real_bpp, code/val.py:161-175, code/bit_counter.py:13-74) on Kodak-shaped inputs: the one
    batched context-model pass that replaces the per-symbol loop, the host range coder, and the sequential decoder."""
    import torch
    from imgcomp_cvpr_b200 import autoencoder, codec, config, probclass, weights
    a, p = config.ae_config('cvpr/low'), config.pc_config('cvpr/res_shallow')
    W = weights.synthetic_weights(a.num_chan_bn, a.num_centers, p.arch_param__k, a.arch_param_B)
    ae = autoencoder.get_network_cls(a)(a, weights=W, mode='exact')
    pc = probclass.get_network_cls(p)(p, num_centers=a.num_centers, weights=W)
    x = weights.synthetic_images(n_images, 768, 512, seed=3)
    imgs = [np.transpose(xi, (1, 2, 0)) for xi in x]
    threads = min(n_images, os.cpu_count() or 1)
    xd = torch.from_numpy(x).cuda()
    sym = ae.encode(xd, is_training=False).symbols.clone()
    centers = ae.centers_tensor()
    tables_ms = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pc.freqs(sym, centers, codec=True)
        e1.record()
        torch.cuda.synchronize()
        tables_ms.append(e0.elapsed_time(e1))
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blobs = codec.compress(imgs, ae, pc, batch_size=n_images, threads=threads)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        rec = codec.decompress(blobs, ae, pc, batch_size=n_images)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        cur = ((t1 - t0) * 1e3, (t2 - t1) * 1e3)
        best = cur if best is None or sum(cur) < sum(best) else best
    # what val.py reconstructs from the encoder's own symbols must be what comes back from the stream alone
    ae.decode(centers[sym], is_training=False)
    same = all(np.array_equal(rec[i], np.transpose(ae.extra['x_out_u8'][i].cpu().numpy(), (1, 2, 0))) for i in range(n_images))
    items = [codec.unpack(b) for b in blobs]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pc.decode_streams([it['stream'] for it in items], [it['first_sym'] for it in items], (32, 96, 64), centers)
    e1.record()
    torch.cuda.synchronize()
    coded_bpp = float(np.mean([8.0 * len(it['stream']) / (768 * 512) for it in items]))
    return {'workload': 'cfg5: %d Kodak-shaped images (768x512), cvpr/low + res_shallow, real bitstreams' % n_images,
            'symbols_per_image': int(sym[0].numel()), 'host_coder_threads': threads, 'coded_bpp': coded_bpp,
            'tables_ms_all_images': min(tables_ms), 'tables_ms_per_image': min(tables_ms) / n_images,
            'compress_ms_per_image': best[0] / n_images, 'decompress_ms_per_image': best[1] / n_images,
            'sequential_decode_kernel_ms_all_images': e0.elapsed_time(e1), 'round_trip_bit_identical': bool(same),
            'reference': 'README.md:65-66: ~350 s encode + ~200 s decode per Kodak image (one sess.run per symbol)'}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='kodak24', choices=sorted(WORKLOADS))
    ap.add_argument('--mode', default=os.environ.get('IC_BENCH_MODE', 'exact'), choices=['fp32', 'exact', 'fast'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the sub-records of the other BASELINE.json configs (N = 1 runs only)')
    ap.add_argument('--with-decode', action='store_true', help='also time ae.decode(qhard) (configs[1] lists it; not part of the metric)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl != 'reference' and not os.environ.get('IC_BENCH_ALLOW_SHORT'):
        args.warmup = 3            # timing rule: at least three untimed warm-up steps (reported in the JSON line)
    args.steps = max(args.steps, 1)
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from imgcomp_cvpr_b200 import weights
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert world == args.gpus, 'launch with torchrun --nproc-per-node %d' % args.gpus
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl')
    res, ctx = measure_encode_pc(args.workload, args.mode, args.steps, args.warmup, world, rank, local,
                                 with_parity=not args.no_parity)
    decode_ms = None
    if args.with_decode:
        ae = ctx['ae']
        enc = ae.encode(ctx['x_dev'], is_training=False)
        qh = enc.qhard.clone()
        for _ in range(2):
            ae.decode(qh, is_training=False)
        ctx['barrier']()
        decode_ms = ctx['timed'](lambda: ae.decode(qh, is_training=False), args.steps)
        ctx['barrier']()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ae_name, N, H, Wd = WORKLOADS[args.workload]
    W, C = ctx['W'], ctx['a'].num_chan_bn
    del ctx
    torch.cuda.empty_cache()
    out = {
        'metric': 'MPix/s encode+probclass fwd', 'value': res['value'], 'unit': 'MPix/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': res['ms_per_step'], 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'fp32': 'f32', 'exact': 'f16x3 (fp32-class, fp32 accumulate)', 'fast': 'f16'}[args.mode],
        'data': 'synthetic',
    }
    for k in ('config', 'e2e', 'gpu_launches', 'clocks', 'roofline', 'kernel_ms_per_step', 'parity'):
        out[k] = res[k]
    out['decode'] = None if decode_ms is None else {'ms_per_step': decode_ms, 'MPix_per_s': world * N * H * Wd / (decode_ms * 1e-3) / 1e6}
    if world == 1 and not args.no_extras:
        # the other BASELINE.json configs, each with its own numbers (N = 1 only: they are not part of the scaling run)
        # (a failure in one of them must not cost the headline line: it is recorded in place of the sub-record)
        k = max(3, min(args.steps, 5))

        def headline():
            h, _ = measure_encode_pc('b64_512', args.mode, k, 3, 1, 0, local, with_parity=not args.no_parity, sample_clocks=False)
            return {'note': 'north-star target shape B=64 512x512 cvpr/low (BASELINE.json north_star)', 'steps': k,
                    'value': h['value'], 'unit': 'MPix/s', 'ms_per_step': h['ms_per_step'], 'e2e': h['e2e'],
                    'roofline': h['roofline'], 'kernel_ms_per_step': h['kernel_ms_per_step'], 'parity': h['parity']}

        extras = []
        if args.workload != 'b64_512':
            extras.append(('headline', headline))
        if args.mode == 'exact':
            extras += [('train_step', lambda: measure_train_step(k)), ('real_bpp', measure_real_bpp)]
        for name, fn in extras:
            try:
                out[name] = fn()
            except Exception as e:      # noqa: BLE001 -- reported, not swallowed
                out[name] = {'error': '%s: %s' % (type(e).__name__, e)}
                print('bench.py: sub-record %s failed: %r' % (name, e), file=sys.stderr)
            try:
                torch.cuda.empty_cache()
            except Exception:           # noqa: BLE001 -- a poisoned context has already been reported above
                pass
    if not args.no_cpu_baseline:
        from oracle import imgcomp_oracle as O
        O.set_backend('torch')
        xs = weights.synthetic_images(1, H, Wd, seed=1234)
        cores = pick_cpu_threads(xs, W, C)
        reps, t0 = 0, time.perf_counter()
        while reps < 2 or (time.perf_counter() - t0 < 10 and reps < 20):
            cpu_reference_step(xs, W, C)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        O.set_backend('numpy')
        out['cpu_baseline'] = {'value': H * Wd / dt / 1e6, 'unit': 'MPix/s', 'cores': cores, 'kind': 'port',
                               'sample': '1 of %d images of %dx%d, %d repetitions' % (N, H, Wd, reps)}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
