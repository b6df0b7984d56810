#!/usr/bin/env python
"""bench.py -- pair scores / second of the all-pairs ViT-ED scoring path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic input:
  * N=1 : BASELINE.json configs[1] -- one synthetic 540-piece puzzle (18x30 grid of 64 px pieces, erosion 7 %,
          4-bin puzzle model patch8/64): 291,060 ordered pairs per step;
  * N>1 : BASELINE.json configs[2] -- 1000-piece puzzles (25x40, erosion 14 %), the (puzzle, row) units of the pair
          grid sharded over the ranks with no data-path collective, 2,500 units per GPU (= exactly configs[2] at
          N=8), one NCCL all-gather of the score blocks at the end of the step ("weak" scaling).
`value` is timed on the device with the piece images already resident in HBM; `e2e` goes through the public API with
HOST buffers (pinned), H2D and D2H inside the timed region. `roofline` is measured live: one extra step with a CUDA
event before every launch (engine option PROFILE) gives every kernel family's summed duration; the family with the
largest share of the step is the roofline object (algorithmic bytes or FLOPs of its launches / that time), the others
are listed under `other_kernels`. `cpu_baseline` / `--impl reference` time the reference's algorithm for this
path (evaluation.py:101-107: one-shot fp32 model(images) on stacked pairs, nothing cached) as restated by the oracle,
on the box's host cores, on a bounded sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'pair scores/sec (all-pairs ViT-ED)'
UNIT = 'pairs/s'
UNITS_PER_GPU = int(os.environ.get('VITED_BENCH_UNITS', '2500'))

# algorithmic FLOPs, puzzle model: D=384, N_e=64, N_d=65, L=8, C=4
F_DEC_PAIR = 8 * (28 * 384 * 384 * 65 + 4 * 65 * 65 * 384 + 4 * 65 * 64 * 384) + 2 * 384 * 4
F_ENC_ITEM = 8 * (24 * 384 * 384 * 64 + 4 * 64 * 64 * 384) + 2 * 64 * 192 * 384
F_KV_ITEM = 8 * 4 * 384 * 384 * 64
F_PREP_ITEM = 2 * 64 * 192 * 384


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d['hbm_gbs'], tf_burst=d['bf16_tflops'], tf_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback (B200_PROFILING.md)')


# ----------------------------------------------------------------------------------------------- workloads
def puzzle_pieces(rows, cols, erosion, seed):
    """Synthetic puzzle -> [rows*cols, 3, 64, 64] fp32 through the reference's a1-a3 steps (pieces.py)."""
    from vited_b200 import pieces, synthetic
    img = synthetic.synthetic_puzzle_image(rows, cols, piece=64, seed=seed)
    lab, grid_size = pieces.make_pieces_lab(img, 64, erosion)
    assert grid_size == (rows, cols)
    return pieces.pieces_to_batch(lab, 64)


def build_model():
    import vited_b200
    from vited_b200 import synthetic
    model = vited_b200.build_model(vited_b200.get_config('puzzle'))
    model.load_state_dict(synthetic.synthetic_state_dict(model, seed=0), strict=True)
    return model


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; falls back to nvidia-smi)."""
    REASONS = {0x4: 'sw_power_cap', 0x8: 'hw_slowdown', 0x20: 'sw_thermal_slowdown', 0x40: 'hw_thermal_slowdown',
               0x80: 'hw_power_brake_slowdown'}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            try:
                uuid = torch.cuda.get_device_properties(device_index).uuid
                self._h = pynvml.nvmlDeviceGetHandleByUUID(('GPU-' + str(uuid)).encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                if self._h is not None:
                    nv = self._nv
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    try:
                        bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                    except Exception:
                        bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                    for bit, name in self.REASONS.items():
                        if bits & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._stop.clear()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


def dist_setup(n_gpus):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        torch.cuda.set_device(0)
    if world != n_gpus:
        raise SystemExit(f'--gpus {n_gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {n_gpus}')
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_rate(budget_s, steps=1, warmup=0, batch=128):
    """The reference's puzzle evaluation forward (evaluation.py:101-107): fp32 one-shot model(images) on stacked
    pairs in batches of 128 (README.md:40), nothing cached -- restated by the oracle, all host threads.
    Returns (pairs/s, cores, sample description, per-step seconds)."""
    from oracle import vited_oracle as orc
    from vited_b200 import synthetic, grid
    import vited_b200
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = vited_b200.build_model(vited_b200.get_config('puzzle'))
    sd = synthetic.synthetic_state_dict(model, seed=0)
    del model
    imgs = puzzle_pieces(4, 8, 0.07, seed=0)  # 32 pieces of the same kind as configs[1]
    pairs = grid.ordered_pairs(imgs.shape[0])

    def run(n_batches):
        done = 0
        for b in range(n_batches):
            sub = pairs[(b * batch) % (len(pairs) - batch):][:batch]
            x = torch.stack([imgs[sub[:, 0]], imgs[sub[:, 1]]], dim=1)
            orc.forward(sd, 12, x)
            done += len(sub)
        return done

    t0 = time.perf_counter()
    run(1)
    probe = time.perf_counter() - t0  # includes first-touch; calibrates the sample size
    per_step_budget = max(budget_s / max(steps + warmup, 1), probe)
    n_batches = max(1, min(8, int(per_step_budget / max(probe, 1e-3))))
    for _ in range(warmup):
        run(n_batches)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        n = run(n_batches)
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    sample = f'{n_batches} batch(es) of {batch} stacked pairs of 64px pieces per step, one-shot fp32 forward (4.43 GFLOP/pair)'
    return n / sec, cores, sample, sec


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    rate, cores, sample, sec = cpu_reference_rate(budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args.gpus),
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
        'note': 'reference model file needs timm (absent); the oracle port of its forward is timed on host cores',
    }
    print(json.dumps(line))


def workload_config(n_gpus):
    if n_gpus == 1:
        return {'workload': 'configs[1]: puzzle all-pairs scoring, one synthetic 540-piece puzzle (18x30) at 64px, '
                            'erosion 7%, 4-bin patch8 model, 291060 ordered pairs per step',
                'pairs_per_step': 540 * 539, 'l2': 'per-step working set (~5 GB of activations per 8065-pair chunk) >> 126 MB L2'}
    return {'workload': f'configs[2]: 1000-piece puzzles (25x40) at 64px, erosion 14%, {UNITS_PER_GPU} (puzzle,row) units '
                        f'per GPU x {n_gpus} GPUs, one NCCL all-gather of the score blocks per step',
            'pairs_per_step': n_gpus * UNITS_PER_GPU * 999, 'parallelism': f'grid rows sharded x{n_gpus}',
            'l2': 'per-step working set >> 126 MB L2'}


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import vited_b200
    from vited_b200 import grid
    # stdout carries exactly ONE JSON line: anything libraries print meanwhile (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device('cuda', local if world > 1 else 0)
    peaks = load_peaks()
    model = build_model().to(dev).eval()

    if world == 1:
        host_images = [puzzle_pieces(18, 30, 0.07, seed=0).pin_memory()]
        row_ranges = [(0, 540)]
        n_pairs_rank = 540 * 539
    else:
        lo_u, hi_u = rank * UNITS_PER_GPU, (rank + 1) * UNITS_PER_GPU
        host_images, row_ranges = [], []
        for pz in range(lo_u // 1000, (hi_u - 1) // 1000 + 1):
            a, b = max(lo_u, pz * 1000) - pz * 1000, min(hi_u, (pz + 1) * 1000) - pz * 1000
            host_images.append(puzzle_pieces(25, 40, 0.14, seed=pz).pin_memory())
            row_ranges.append((a, b))
        n_pairs_rank = UNITS_PER_GPU * 999
    n_pairs_job = n_pairs_rank * world
    dev_images = [h.to(dev) for h in host_images]
    n_items = dev_images[0].shape[0]
    out_block = torch.zeros((sum(b - a for a, b in row_ranges), n_items, 4), dtype=torch.float32, device=dev)
    gathered = torch.empty((world,) + tuple(out_block.shape), dtype=torch.float32, device=dev) if world > 1 else None
    host_out = torch.empty(out_block.shape, dtype=torch.float32).pin_memory()

    def step(images):
        r0 = 0
        for img, (a, b) in zip(images, row_ranges):
            model.score_grid(img, vited_b200.GRID_ORDERED_OFFDIAG, a, b, out=out_block[r0:r0 + (b - a)])
            r0 += b - a
        if world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(gathered, out_block)
        return out_block

    def step_e2e():
        imgs = [h.to(dev, non_blocking=True) for h in host_images]
        res = step(imgs)
        host_out.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        step(dev_images)
    barrier(world)
    clocks = ClockSampler(dev.index or 0)
    launches0 = model.launch_count()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(dev_images)
    e1.record()
    torch.cuda.synchronize()
    clock_info = clocks.stop()
    barrier(world)
    launches = model.launch_count() - launches0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps, world)
    value = n_pairs_job / (ms_step / 1e3)

    # ---- end to end through the public API with host buffers
    for _ in range(2):
        step_e2e()
    barrier(world)
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step_e2e()
    t1.record()
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks(t0.elapsed_time(t1) / args.steps, world)
    e2e_value = n_pairs_job / (ms_e2e / 1e3)
    h2d = int(sum(h.numel() * 4 for h in host_images))
    d2h = int(host_out.numel() * 4)

    # ---- per-kernel timing of one more step (roofline of the dominant kernel)
    model.set_option(vited_b200.OPT_PROFILE, 1)
    step(dev_images)
    prof = model.profile_read()
    model.set_option(vited_b200.OPT_PROFILE, 0)
    total_ms = sum(v['ms'] for v in prof.values()) or 1.0
    # kernel families of the step; the roofline object describes the one with the largest share of the step
    #   gemm_ln : gemm_ln_pair_kernel (Linear + residual + LayerNorm fused, N = 384): bound by HBM (fp32 residual in/out)
    #   gemm    : gemm_tc_pair_kernel / gemm_tc_kernel (tcgen05 GEMMs with bias / GELU epilogue): tensor bound
    #   attn    : attn_p64_kernel (tcgen05 attention, 65-token sequences): HBM bound
    # `traffic` = dram__bytes_read.sum + dram__bytes_write.sum per launch. ncu --set full was captured at 262,080 rows per
    # launch (profiles/r01b_ncu_full_ops_raw.csv): measured DRAM bytes / algorithmic bytes, launch-weighted over the
    # family's shapes, was 0.964 (gemm_ln), 0.935 (gemm), 0.94 (attn) -- no re-reads; the ratio is applied to the
    # algorithmic bytes per launch of THIS run (the launches are per-row uniform, the default chunk is larger now).
    fams = {
        'gemm_ln': dict(match=lambda k: k.startswith('gemm_ln_'), bound='hbm', traffic_ratio=0.964,
                        kernel='gemm_ln_pair_kernel (tcgen05 Linear + residual + LayerNorm, full-row epilogue out of TMEM)'),
        'gemm': dict(match=lambda k: k.startswith('gemm_n'), bound='tensor', traffic_ratio=0.935,
                     kernel='gemm_tc_pair_kernel (tcgen05/TMEM/TMA cta_group::2 GEMM, bias / GELU epilogue)'),
        'attn': dict(match=lambda k: k in ('attn_self', 'attn_cross'), bound='hbm', traffic_ratio=0.94,
                     kernel='attn_p64_kernel (tcgen05 attention, S/P/O in TMEM)'),
    }

    def family_roofline(name):
        f = fams[name]
        sel = {k: v for k, v in prof.items() if f['match'](k)}
        ms = sum(v['ms'] for v in sel.values())
        n = sum(v['launches'] for v in sel.values())
        if ms <= 0:
            return None
        if f['bound'] == 'tensor':
            ach = sum(v['flops'] for v in sel.values()) / (ms / 1e3) / 1e12
            peak, unit = peaks['tf_sustained'], 'TFLOP/s'
        else:
            ach = sum(v['bytes'] for v in sel.values()) / (ms / 1e3) / 1e9
            peak, unit = peaks['hbm_gbs'], 'GB/s'
        return {'bound': f['bound'], 'kernel': f['kernel'], 'achieved': ach, 'peak': peak, 'unit': unit, 'frac': ach / peak,
                'peak_source': peaks['source'] + (', sustained dense 16-bit figure (measured by the driver with bf16; fp16 operands run at the same rate; kernel timed inside a long step)' if f['bound'] == 'tensor' else ''),
                'traffic': f['traffic_ratio'] * sum(v['bytes'] for v in sel.values()) / max(n, 1) if world == 1 else None,
                'launches_per_step': n, 'avg_launch_ms': ms / max(n, 1),
                'algorithmic_per_launch': (sum(v['flops'] for v in sel.values()) if f['bound'] == 'tensor'
                                           else sum(v['bytes'] for v in sel.values())) / max(n, 1),
                'share_of_step': ms / total_ms}

    fam_lines = {k: family_roofline(k) for k in fams}
    fam_lines = {k: v for k, v in fam_lines.items() if v}
    top = max(fam_lines, key=lambda k: fam_lines[k]['share_of_step'])
    roofline = dict(fam_lines[top])
    roofline['other_kernels'] = {k: v for k, v in fam_lines.items() if k != top}
    roofline['classes'] = {k: {'ms': round(v['ms'], 3), 'share': round(v['ms'] / total_ms, 4),
                               'tflops': round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 1),
                               'gbs': round(v['bytes'] / max(v['ms'], 1e-9) / 1e6, 1), 'launches': v['launches']}
                           for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}
    # whole-step algorithmic work against the tensor roofline
    n_items_total = sum(img.shape[0] for img in dev_images)
    n_ctx = sum(b - a for a, b in row_ranges)
    f_step = n_pairs_rank * F_DEC_PAIR + n_items_total * F_PREP_ITEM + n_ctx * (F_ENC_ITEM + F_KV_ITEM)
    step_frac = f_step / (ms_step / 1e3) / 1e12 / peaks['tf_sustained']
    # FLOPs actually launched (layer-0 self-attention runs once per item, the last layer runs on class-token rows
    # only): sum of the per-launch algorithmic flops the engine attached to the profiled step
    f_exec = sum(v['flops'] for v in prof.values())
    step_frac_exec = f_exec / (ms_step / 1e3) / 1e12 / peaks['tf_sustained']

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': vited_b200.ACT_NAME,
            'data': 'synthetic', 'config': workload_config(world),
            'clocks': clock_info,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'ms_per_step': ms_e2e},
            'gpu_launches': int(launches),
            'roofline': roofline,
            'step_tensor_frac': step_frac,
            'step_tensor_frac_executed': step_frac_exec,
            'flops_note': 'step_tensor_frac uses SURVEY 8d nominal work (2.250 GFLOP/pair, all 8 decoder layers on all 65 '
                          'rows); _executed counts only launched flops (layer-0 self-attention cached per item, last '
                          'layer pruned to the class-token row)',
        }
        if world == 1 and not args.no_cpu:
            rate, cores, sample, _ = cpu_reference_rate(budget_s=20.0, steps=1, warmup=0)
            line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def run_hisfrag(args):
    """Side measurement (not the driver's line): BASELINE.json configs[3] model (patch16/512, 12+12 layers) on a
    small all-pairs grid of `--items` synthetic fragments on ONE GPU; reports pairs/s and the tensor-roofline
    fraction with the Hisfrag FLOP model of SURVEY 8d (89.50 GFLOP / pair, 63.42 + 7.25 + 0.60 GFLOP / item)."""
    import vited_b200
    from vited_b200 import grid, synthetic
    torch.cuda.set_device(0)
    peaks = load_peaks()
    model = vited_b200.build_model(vited_b200.get_config('hisfrag'))
    model.load_state_dict(synthetic.synthetic_state_dict(model, seed=0), strict=True)
    model = model.cuda().eval()
    n = args.items
    images = synthetic.synthetic_images(n, 512, seed=1000).cuda()
    n_pairs = n * (n + 1) // 2
    for _ in range(max(args.warmup, 1)):
        grid.score_fragments(model, images)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sim = grid.score_fragments(model, images)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    model.set_option(vited_b200.OPT_PROFILE, 1)
    grid.score_fragments(model, images)
    prof = model.profile_read()
    model.set_option(vited_b200.OPT_PROFILE, 0)
    total = sum(v['ms'] for v in prof.values()) or 1.0
    flops = n_pairs * 89.50e9 + n * (63.42e9 + 7.248e9 + 0.604e9)
    print(json.dumps({
        'metric': METRIC, 'value': n_pairs / (ms / 1e3), 'unit': UNIT, 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms, 'dtype': vited_b200.ACT_NAME, 'data': 'synthetic',
        'config': {'workload': f'configs[3] model (Hisfrag20 patch16 512px), {n} synthetic fragments, {n_pairs} pairs (a<=b), 1 GPU'},
        'step_tensor_frac': flops / (ms / 1e3) / 1e12 / peaks['tf_sustained'],
        'classes': {k: {'ms': round(v['ms'], 3), 'share': round(v['ms'] / total, 4),
                        'tflops': round(v['flops'] / max(v['ms'], 1e-9) / 1e9, 1), 'launches': v['launches']}
                    for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg (profiling runs)')
    ap.add_argument('--workload', default='puzzle', choices=['puzzle', 'hisfrag'],
                    help="'hisfrag' = side measurement of the Hisfrag20 model on a small grid (1 GPU)")
    ap.add_argument('--items', type=int, default=48, help='fragments for --workload hisfrag')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = max(args.warmup, 0)
    if args.impl == 'reference':
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit('bench.py: no CUDA device; the scoring path has no CPU fallback (use --impl reference)')
        if args.workload == 'hisfrag':
            run_hisfrag(args)
        else:
            run_ours(args)


if __name__ == '__main__':
    main()
