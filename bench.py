#!/usr/bin/env python
"""bench.py -- mini-batches/sec of the GRU4Rec session-parallel training step on synthetic RSC15-shaped sessions.

Contract: python bench.py --gpus N --steps K --warmup W   (torchrun for N>1) prints ONE JSON line.
  value   : whole-job mini-batches/s with the schedule window, sample store and parameters resident in HBM
  e2e     : the same metric through the reference-facing call (host schedule arrays -> H2D -> steps -> D2H costs)
  roofline: dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline: the NumPy oracle (port of the reference; Theano is not installable) on the host cores, bounded sample
--impl reference : times that CPU port alone (rank 0 only), same metric / config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# workload = BASELINE.json configs[1]: RSC15 1-layer GRU(100) BPR-max n_sample=2048 batch=32 (param_samples/rsc15_bpr-max.py)
WORKLOAD = dict(name='rsc15_bprmax_gru100_b32_ns2048', n_items=37483,
                model=dict(layers=[100], loss='bpr-max', final_act='elu-0.5', hidden_act='tanh', batch_size=32, dropout_p_embed=0.0,
                           dropout_p_hidden=0.0, learning_rate=0.2, momentum=0.3, sample_alpha=0.0, n_sample=2048, bpreg=1.0,
                           constrained_embedding=False),
                sample_store=10000000)
ALGO_BYTES_PER_STEP = 6041792
# the reference's published mini-batches/s for this configuration (BPR-max, B=32, GRU(100), n_sample=2048, with
# momentum) -- read off its training-time chart, measured on an A30 (img/training_time_bprmax_batch_size.png, README.md:302)
BASELINE_PUBLISHED_MBS = 1235.0


def algo_bytes_lossgrad(N, L, mom=True):
    """sparse Adagrad(+momentum) of the N gathered Wy rows + By: (param, acc[, vel]) read + write."""
    T = 6 if mom else 4
    return T * 4 * (N * L + N)


class ClockSampler(threading.Thread):
    def __init__(self, gpu_index=0):
        threading.Thread.__init__(self, daemon=True)
        self.rows = []
        self.stop_flag = False
        self.gpu_index = gpu_index

    def run(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            p = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu_index), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '100'],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return
        while not self.stop_flag:
            line = p.stdout.readline()
            if not line:
                break
            self.rows.append([x.strip() for x in line.split(',')])
        p.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(np.max(mx)), 'reasons': sorted(reasons), 'samples': len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def build_workload(n_steps_needed, seed=0):
    from gru4rec_b200.synth import make_session_arrays
    B = WORKLOAD['model']['batch_size']
    n_events = max(int((n_steps_needed + 64) * B * 1.6) + 20000, 4 * WORKLOAD["n_items"])
    return make_session_arrays(WORKLOAD['n_items'], n_events, seed=seed)


def oracle_steps_per_sec(items, offset, order, supports, n_warm, n_steps, budget_s):
    """The NumPy restatement of the reference step (oracle/) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import gru4rec_oracle as orc
    mk = dict(WORKLOAD['model'])
    m = orc.OracleGRU4Rec(**mk)
    m.init(WORKLOAD['n_items'])
    P = orc.sampling_cdf(supports, mk['sample_alpha']).astype(np.float32)
    rs = np.random.RandomState(1)
    B = mk['batch_size']
    # literal schedule restatement on a prefix of the data (the schedule itself is outside the timed step)
    n_sess = int(np.searchsorted(offset, (n_warm + n_steps + 8) * B * 3))
    n_sess = max(min(n_sess, len(offset) - 1), B + 1)
    steps = orc.build_train_schedule(items, offset[:n_sess + 1], order[:n_sess], B, mk['n_sample'])
    steps = steps[:n_warm + n_steps]
    t_start = time.time()
    done = 0
    t0 = None
    for k, st in enumerate(steps):
        if k == n_warm:
            t0 = time.time()
        smp = orc.searchsorted_k2(P, rs.rand(mk['n_sample']).astype(np.float32))
        m.train_step(st['X'], st['Y'], st['R'], samples=smp, slots=st['slots'])
        if k >= n_warm:
            done += 1
            if time.time() - t_start > budget_s:
                break
    dt = time.time() - t0
    try:
        from threadpoolctl import threadpool_info
        cores = max([p.get('num_threads', 1) for p in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count()
    return done / dt, done, cores


def run_reference(args, rank):
    if rank != 0:
        return
    n = args.steps
    items, offset, order, supports = build_workload(min(n, 4000) + args.warmup)
    v, done, cores = oracle_steps_per_sec(items, offset, order, supports, args.warmup, n, budget_s=150.0)
    out = {
        'metric': 'mini-batches/sec', 'value': v, 'unit': 'mb/s', 'n_gpus': args.gpus, 'steps': done, 'warmup': args.warmup,
        'ms_per_step': 1000.0 / v, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': v / BASELINE_PUBLISHED_MBS, 'dtype': 'f32', 'data': 'synthetic',
        'impl': 'reference',
        'config': {'workload': WORKLOAD['name'], 'n_items': WORKLOAD['n_items'], 'global_batch': WORKLOAD['model']['batch_size'],
                   'note': 'reference CPU path = NumPy restatement of gru4rec.py (oracle/); Theano is not installable offline'},
        'cpu_baseline': {'value': v, 'unit': 'mb/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d timed mini-batches of the same workload after %d warm-up (time-bounded)' % (done, args.warmup)},
        'e2e': {'value': v, 'unit': 'mb/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4000)
    ap.add_argument('--warmup', type=int, default=200)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--step-mode', type=int, default=2, help='0 per-phase kernels (CUDA graph), 1 persistent kernel, 2 role-specialised persistent kernel (default), 3 = 2 with the GRU phases on one thread-block cluster')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    import torch
    from gru4rec_b200 import _lib
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    K, W = args.steps, max(args.warmup, 3)
    make_cfg = _lib.make_config
    mk = dict(WORKLOAD['model'])
    cfg = make_cfg(WORKLOAD['n_items'], mk, sample_store=WORKLOAD['sample_store'], eval_lanes=0,
                   max_resident_steps=min(max(K, W), WORKLOAD['sample_store'] // mk['n_sample']) + 8, step_mode=args.step_mode, world_size=world, rank=rank)
    eng = _lib.Engine(cfg, device=local_rank)
    if world > 1:
        eng.init_multi_gpu(dist)
    # parameters: the reference's initialisation (gru4rec.py:254-294); data: synthetic RSC15-shaped sessions, disjoint per rank
    import gru4rec as g4
    gru = g4.GRU4Rec(**mk)
    gru.n_items = WORKLOAD['n_items']
    host = gru._init_host_weights()
    for name, w in host.items():
        eng.set(name, w)
    items, offset, order, supports = build_workload(2 * (K + W), seed=rank)
    P = supports.astype(np.float64) ** mk['sample_alpha']
    P = P.cumsum() / P.sum(); P[-1] = 1
    eng.set_sampling_cdf(P.astype(np.float32))
    eng.generate_samples()
    sched = _lib.Schedule(items, offset, order, mk['batch_size'], mk['n_sample'], mode=0)
    assert sched.n_steps >= 2 * (K + W), 'synthetic workload too small'
    B = mk['batch_size']
    N = B + mk['n_sample']

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank); clocks.start()
    time.sleep(0.3)
    h2d = B * (4 + 4 + 4 + 1) + 12
    first = W + K
    if world == 1:
        # ---- device-resident arm: warm-up, then K timed steps from uploaded windows.  A window never crosses a refill of the
        # negative-sample store (4882 mini-batches at the headline shape): uploads and refills happen between the timed
        # windows; `value` sums the CUDA-event times of the windows (the e2e arm below times everything, refills included).
        eng.reset_hidden()
        eng.upload_steps(sched, 0, W)
        eng.run_uploaded(W, want_cost=False)
        gen_len = eng.sample_store_rows()
        cap = int(cfg.max_resident_steps)
        launches0 = eng.kernel_launches()
        barrier()
        t0 = time.time()
        dev_ms, done, cost_parts = 0.0, 0, []
        while done < K:
            if eng.get_sample_pointer() >= gen_len:
                eng.generate_samples()
            n = min(K - done, gen_len - eng.get_sample_pointer(), cap)
            eng.upload_steps(sched, W + done, n)
            c, ms = eng.run_uploaded(n, want_cost=True)
            dev_ms += ms; done += n; cost_parts.append(c)
        barrier()
        wall = time.time() - t0
        costs = np.concatenate(cost_parts)
        launches = eng.kernel_launches() - launches0
        value = K / (dev_ms / 1000.0)
        # ---- end-to-end arm: host schedule arrays in, costs out, every window (H2D + plan + steps + D2H inside the timing)
        barrier()
        t0 = time.time()
        c2 = eng.train_steps(sched, first, K)
        barrier()
        e2e_s = time.time() - t0
    else:
        # ---- N ranks in lock step: one merged update per mini-batch (NCCL all-gather of row gradients + all-reduce of the
        # dense gradients inside g4r_train_steps).  The call takes HOST schedule arrays, so this IS the end-to-end path;
        # timed with barrier + synchronize on both sides, max over ranks.
        eng.reset_hidden()
        eng.train_steps(sched, 0, W)
        launches0 = eng.kernel_launches()
        barrier()
        t0 = time.time()
        costs = eng.train_steps(sched, W, K)
        barrier()
        wall = time.time() - t0
        launches = eng.kernel_launches() - launches0
        t = torch.tensor([wall], device='cuda'); dist.all_reduce(t, op=dist.ReduceOp.MAX); wall = float(t.item())
        dev_ms = wall * 1000.0
        value = world * K / wall
        e2e_s = wall
    assert np.isfinite(costs).all(), 'non-finite cost in the timed region'
    clocks.stop_flag = True
    e2e_value = world * K / e2e_s
    # ---- per-kernel roofline from CUDA events around every launch of one more pass over a short window
    prof_n = min(K, 512)
    if world == 1:
        if eng.get_sample_pointer() + prof_n > eng.sample_store_rows():
            eng.generate_samples()
        eng.upload_steps(sched, first + K, prof_n)
        prof = eng.profile_uploaded()
    else:
        prof = None      # the per-kernel roofline is a single-GPU measurement (N=1 run of this same script)
    peak, peak_src = peak_hbm()
    lg_bytes = algo_bytes_lossgrad(N, mk['layers'][-1], mk['momentum'] > 0)
    fast_phase = None
    if world == 1 and int(cfg.step_mode) in (2, 3) and eng.fast_windows()[0] > 0:
        # the same update phase INSIDE the production kernel k_fast, from %globaltimer stamps of CTA 0 (slot 2 = row statistics
        # ready, slot 14 = chunk's rows updated): loss gradient + dSy + partial dL/dh + sparse update of the chunk's rows
        eng.persistent_stamps(True)
        if eng.get_sample_pointer() + prof_n > eng.sample_store_rows():
            eng.generate_samples()
        eng.upload_steps(sched, first + K, prof_n)
        eng.run_uploaded(prof_n, want_cost=False)
        st = eng.persistent_stamps(False, prof_n).astype(np.int64)
        seg_us = float(np.mean((st[8:, 14] - st[8:, 2]) / 1000.0))
        fast_phase = {'us': seg_us, 'achieved_GBs': lg_bytes / (seg_us * 1e-6) / 1e9, 'frac': lg_bytes / (seg_us * 1e-6) / 1e9 / peak,
                      'note': 'k_fast: statistics-ready -> rows-updated segment of a chunk CTA (globaltimer), same algorithmic bytes'}
    if prof is not None:
        dom_name = max(prof, key=lambda k: prof[k][0])
        lg_ms, lg_n = prof['lossgrad_update']
        achieved = lg_bytes / (lg_ms / lg_n * 1e-3) / 1e9
        frac, us_launch = achieved / peak, lg_ms / lg_n * 1000.0
        phase_us = {k: round(v[0] / v[1] * 1000.0, 3) for k, v in prof.items()}
    else:
        dom_name, achieved, frac, us_launch, phase_us = None, None, None, None, None
    out = {
        'metric': 'mini-batches/sec', 'value': value, 'unit': 'mb/s', 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': dev_ms / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': value / BASELINE_PUBLISHED_MBS,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD['name'], 'n_items': WORKLOAD['n_items'], 'global_batch': B * world, 'n_sample': mk['n_sample'],
                   'layers': mk['layers'], 'params': 'param_samples/rsc15_bpr-max.py', 'parallelism': ('dp%d: replicated parameters, NCCL all-gather of row gradients + all-reduce of dense gradients per mini-batch, identical merged update on every rank' % world) if world > 1 else 'dp1',
                   'l2': 'inputs larger than L2: item tables + Adagrad/momentum state = 180 MB, rows touched change every step (no flush between steps; ncu shows the uniformly sampled Wy rows, 45 MB with their state, staying L2-resident in steady state: 0.23 MB DRAM traffic per step)',
                   'step_mode': int(cfg.step_mode), 'fast_windows': list(eng.fast_windows()), 'events_per_sec': value * B,
                   'vs_baseline_source': 'BASELINE.md: ~1235 mb/s published by the reference for this configuration on an A30 (chart reading)'},
        'e2e': {'value': e2e_value, 'unit': 'mb/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'clocks': clocks.summary(),
        'roofline': {'bound': 'hbm', 'kernel': 'k_lossgrad (loss gradient + sparse Adagrad/momentum update of Wy/By rows)',
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': frac, 'traffic': 3812352 if world == 1 else None,
                     'peak_source': peak_src, 'algorithmic_bytes_per_launch': lg_bytes, 'us_per_launch': us_launch, 'traffic_source': 'ncu --set full dram read+write of k_lossgrad, profiles/r1_ncu_full_k_lossgrad.txt',
                     'dominant_phase_by_time': dom_name, 'phase_us': phase_us, 'k_fast_update_phase': fast_phase,
                     'k_fast_dram_traffic_per_step': {'bytes': 228705, 'source': 'ncu --set full of the timed window (2000 mini-batches in one launch): 202.3 MB read + 255.1 MB written, profiles/r1_ncu_full_k_fast_steady.txt; the touched rows stay in the 126 MB L2'},
                     'whole_step': {'algorithmic_bytes': ALGO_BYTES_PER_STEP, 'achieved': ALGO_BYTES_PER_STEP * (value / world) / 1e9,
                                    'frac': ALGO_BYTES_PER_STEP * (value / world) / 1e9 / peak,
                                    'note': 'latency-bound: dependent phases per mini-batch, working set near L2 size'}},
        'wall_s_timed_region': wall,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, done, cores = oracle_steps_per_sec(items, offset, order, supports, 10, 2000, budget_s=20.0)
        out['cpu_baseline'] = {'value': v, 'unit': 'mb/s', 'cores': cores, 'kind': 'port',
                               'sample': '%d mini-batches of the same workload through the NumPy oracle (~20 s)' % done}
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
