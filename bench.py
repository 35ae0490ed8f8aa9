#!/usr/bin/env python
"""bench.py — samples/sec per fit iteration (forward + gradient + all-reduce + penalty/Adam step)
of the CP tensor-regression hot path, and its fraction of the HBM roofline (2 passes x bytes(X)).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

N > 1 is launched by torchrun (one rank per GPU, NCCL); the sample axis is sharded (each rank
holds its own slice of X resident in HBM: weak scaling) and the packed gradient sums are
all-reduced once per iteration.  Rank 0 prints ONE JSON line.

Workloads (BASELINE.json configs; the default is configs[1], the one the metric is quoted on):
  cfg1  standard   X (2000, 20,30,40)      R=5   fp32
  cfg2  standard   X (200000, 64,64,32)    R=8   fp32   (104.9 GB of X per GPU)
  cfg3  multinomial X (62500/GPU, 100,50,20) C=10 R=6 fp32   (500000 over 8 GPUs)
  cfg4  standard   X (100000, 16,16,16,32) R=12  fp64
  cfg5  multinomial X (312500/GPU, 100,50,20) C=4 R=4 fp32   (~1 TB over 8 GPUs)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    #        kind   N per GPU  dims                R   C   dtype
    'cfg1': ('std', 2000, (20, 30, 40), 5, 0, torch.float32),
    'cfg2': ('std', 200000, (64, 64, 32), 8, 0, torch.float32),
    'cfg3': ('mn', 62500, (100, 50, 20), 6, 10, torch.float32),
    'cfg4': ('std', 100000, (16, 16, 16, 32), 12, 0, torch.float64),
    'cfg5': ('mn', 312500, (100, 50, 20), 4, 4, torch.float32),
    # diagnostic only (not a BASELINE config): the standard model on the cfg3 sample shape
    'dbg3': ('std', 62500, (100, 50, 20), 6, 0, torch.float32),
}
DESCR = {
    'cfg1': 'standard CP regression, X (N=2000, 20,30,40), rank 5, fp32',
    'cfg2': 'standard CP regression, X (N=200000, 64,64,32), rank 8, fp32',
    'cfg3': 'multinomial CP regression, X (N=62500 per GPU, 100,50,20), n_classes=10, rank 6, fp32',
    'cfg4': '5-mode standard CP regression, X (N=100000, 16,16,16,32), rank 12, fp64',
    'cfg5': 'multinomial CP regression, X (N=312500 per GPU, 100,50,20), n_classes=4, rank 4, fp32',
    'dbg3': 'diagnostic: standard CP regression on the cfg3 sample shape, X (N=62500, 100,50,20), rank 6, fp32',
}
ADAM = {'lr': 0.01, 'amsgrad': True}
LAMBDA = 0.01


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md)'


def bind_to_gpu_numa_node(index):
    """Multi-rank runs: pin this process to the CPUs NVML reports as local to its GPU, so the pinned host
    buffers of the e2e legs are first-touched on the GPU's own NUMA node (eight ranks streaming from the
    wrong socket share one inter-socket link).  Returns the CPU count of the set, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        phys = index
        if vis:
            ids = [v.strip() for v in vis.split(',') if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w in range(words) for b in range(64) if (int(mask[w]) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML in a thread (a sample every
    few ms; the timed region of the default run is ~0.2 s), nvidia-smi -lms as the fallback."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self._stop = index, [], None, None, False

    def _phys_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            ids = [v.strip() for v in vis.split(',') if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.hdl = pynvml.nvmlDeviceGetHandleByIndex(self._phys_index())
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.hdl, pynvml.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self._phys_index()}', f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown,
                'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.hdl, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.hdl))
                self.rows.append((sm, [k for k, b in bits.items() if r & b]))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.nvml is not None:
            self._stop = True
            self.t.join(timeout=2)
            sm = [r[0] for r in self.rows]
            reasons = sorted({k for r in self.rows for k in r[1]})
            return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': self.mx, 'reasons': reasons,
                    'samples': len(sm), 'source': 'nvml'}
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm), 'source': 'nvidia-smi'}


# ----------------------------------------------------------------------------------------------
# synthetic workload, generated on the device in chunks
# ----------------------------------------------------------------------------------------------
def make_device_data(wl, n_local, rank, device):
    kind, _, dims, R, C, dt = WORKLOADS[wl]
    g = torch.Generator(device=device).manual_seed(1234 + int(wl[-1]) + 1000 * rank)   # (last character: config number)
    X = torch.empty((n_local, *dims), dtype=dt, device=device)
    step = max(1, (1 << 28) // int(np.prod(dims)))
    for lo in range(0, n_local, step):
        X[lo:lo + step].normal_(generator=g)
    gc = torch.Generator().manual_seed(4321)
    Fstar = [0.3 * torch.randn((d, R), generator=gc, dtype=dt) for d in list(dims) + ([C] if C else [])]
    return X, Fstar


class CycledHostArray:
    """A (N, ...) host array whose samples repeat a pinned pool — lets the e2e leg stream the full
    N x D bytes per step over PCIe without needing N x D bytes of host RAM (data is synthetic)."""

    def __init__(self, pool, n):
        self.pool, self.n = pool, int(n)
        self.shape = (self.n, *pool.shape[1:])

    def __getitem__(self, sl):
        lo, hi, _ = sl.indices(self.n)
        p = self.pool.shape[0]
        a = lo % p
        if a + (hi - lo) <= p:
            return self.pool[a:a + (hi - lo)]
        return torch.cat([self.pool[a:], self.pool[:(hi - lo) - (p - a)]])


def cpu_baseline(wl, budget_s=12.0, n_s=None):
    """The reference's algorithm (oracle port: dense B + one matmul + autograd, torch CPU, all host
    threads) on a bounded sample of the same workload.  Returns samples/s per fwd+grad iteration."""
    from oracle import tr_oracle as O
    kind, _, dims, R, C, dt = WORKLOADS[wl]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D = int(np.prod(dims))
    if n_s is None:
        n_s = int(max(64, min(2000, (1 << 30) // (D * (4 if dt == torch.float32 else 8)))))
    if kind == 'std':
        X, y, _ = O.synth_std(n_s, dims, R, 1234 + 2, dtype=dt)
        nn = [False] * (len(dims) + 1)
        B0 = O.init_std(dims, R, nn, dtype=dt)
        bias, w = torch.tensor([0.0], dtype=dt), torch.ones(R, dtype=dt)
        fn = lambda: O.std_loss_grad(X, y, B0, bias, w, nn, LAMBDA)  # noqa: E731
    else:
        X, y, _ = O.synth_mn(n_s, dims, R, C, 1234 + 3)
        nn = [False] * (len(dims) + 1)
        B0 = O.init_mn(list(dims) + [C], R, nn)
        w, cw = torch.ones(R), np.ones(C, dtype=np.float32)
        fn = lambda: O.mn_loss_grad(X, y, B0, w, nn, cw, LAMBDA)  # noqa: E731
    for _ in range(3):
        fn()
    ts, t_end = [], time.perf_counter() + budget_s
    while len(ts) < 200 and (time.perf_counter() < t_end or len(ts) < 5):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    med = float(np.median(ts))
    return {'value': n_s / med, 'unit': 'samples/s', 'cores': cores, 'kind': 'port',
            'sample': f'{n_s} samples of the {wl} sample shape {tuple(dims)}, oracle port of the reference '
                      f'(dense B + matmul + autograd, torch CPU {torch.get_num_threads()} threads), '
                      f'median of {len(ts)} fwd+grad iterations',
            'ms_per_iter': med * 1e3}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the host cores
    (oracle port here: the Python reference cannot travel to the GPU box), rank 0 only."""
    if rank != 0:
        return
    wl = args.workload
    kind, _, dims, R, C, dt = WORKLOADS[wl]
    from oracle import tr_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D = int(np.prod(dims))
    n_s = int(max(64, min(2000, (1 << 30) // (D * (4 if dt == torch.float32 else 8)))))
    if kind == 'std':
        X, y, _ = O.synth_std(n_s, dims, R, 1234 + 2, dtype=dt)
        nn = [False] * (len(dims) + 1)
        B = [b.clone().requires_grad_(True) for b in O.init_std(dims, R, nn, dtype=dt)]
        bias = torch.tensor([0.0], dtype=dt, requires_grad=True)
        w = torch.ones(R, dtype=dt)
        opt = torch.optim.Adam(B + [bias], **ADAM)
        loss_fn = torch.nn.MSELoss()

        def step():
            opt.zero_grad()
            loss = loss_fn(O.lin_model(X, B, w, nn, bias), y) + LAMBDA * O.L2_penalty(B)
            loss.backward()
            opt.step()
            return loss.item()
    else:
        X, y, _ = O.synth_mn(n_s, dims, R, C, 1234 + 3)
        nn = [False] * (len(dims) + 1)
        B = [b.clone().requires_grad_(True) for b in O.init_mn(list(dims) + [C], R, nn)]
        w = torch.ones(R)
        opt = torch.optim.Adam(B, **ADAM)
        loss_fn = torch.nn.CrossEntropyLoss(weight=torch.ones(C))

        def step():
            opt.zero_grad()
            loss = loss_fn(O.mn_model(X, B, w, nn), y) + LAMBDA * O.L2_penalty(B)
            loss.backward()
            opt.step()
            return loss.item()
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt_s = time.perf_counter() - t0
    val = n_s * args.steps / dt_s
    sample = (f'{n_s} samples of the {wl} sample shape {tuple(dims)} per step (bounded sample of the workload), '
              f'oracle port of the reference fit_Adam iteration on torch CPU, {torch.get_num_threads()} threads')
    out = {'impl': 'reference', 'metric': 'samples/sec per fit iteration (fwd+grad+step)', 'value': val,
           'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
           'ms_per_step': dt_s / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32' if dt == torch.float32 else 'f64', 'data': 'synthetic',
           'config': {'workload': DESCR[wl], 'sample': sample},
           'cpu_baseline': {'value': val, 'unit': 'samples/s', 'cores': cores, 'kind': 'port', 'sample': sample},
           'e2e': {'value': val, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything else a library prints on fd 1 (e.g. the
    NCCL version banner) was redirected to stderr at start-up."""
    line = (json.dumps(obj) + '\n').encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=sorted(WORKLOADS))
    ap.add_argument('--n-local', type=int, default=0, help='override samples per GPU (debug)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--fused', type=int, default=-1, help='-1 auto, 0 never, 1 always: single-pass cluster kernel (std)')
    ap.add_argument('--flow', type=int, default=-1, help='-1 auto, 0 never, 1 always: single-launch dataflow kernel')
    ap.add_argument('--flow-debug', type=int, default=0, help='timing experiments only (wrong results)')
    ap.add_argument('--flow-window-mb', type=int, default=0, help='dataflow kernel: MB of X kept in flight in L2 (0 = default)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from tensor_regression_b200 import standard_tensor_regression as STR
    from tensor_regression_b200 import multinomial_tensor_regression as MTR
    from tensor_regression_b200 import engine as E

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU path)'
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', rank=rank, world_size=world, device_id=device)
    if args.warmup < 3 and rank == 0:
        print(f'note: --warmup {args.warmup} < 3 (timing rules ask for >= 3)', file=sys.stderr)

    wl = args.workload
    kind, n_local, dims, R, C, dt = WORKLOADS[wl]
    if args.n_local:
        n_local = args.n_local
    D = int(np.prod(dims))
    elt = 4 if dt == torch.float32 else 8
    x_bytes = n_local * D * elt

    # ---- data resident in HBM ------------------------------------------------------------
    X, Fstar = make_device_data(wl, n_local, rank, device)
    sharder = E.ShardedSum() if world > 1 else E.ShardedSum(enabled=False)
    nn = [False] * (len(dims) + 1)
    torch.manual_seed(321)
    if kind == 'std':
        eng = E.Engine(dims, R, 0, dt, device)
        w = torch.ones(R, dtype=dt, device=device)
        theta_star = torch.cat([f.reshape(-1) for f in Fstar] + [torch.tensor([0.1], dtype=dt)]).to(device)
        y = eng.forward_std(X, theta_star, w, 0, 50.0, 1.0)
        y += 0.01 * torch.randn(y.shape, dtype=dt, device=device)
        B0 = STR.make_BcpInit(list(dims), R, nn, scale=1, device='cpu', dtype=dt)
        model = STR.CP_linear_regression((n_local, *dims), dtype=dt, rank=R, non_negative=False, Bcp_init=B0,
                                         device=device, shard_group='world' if world > 1 else None)
        cw = None
    else:
        eng = E.Engine(dims, R, C, torch.float32, device)
        w = torch.ones(R, device=device)
        theta_star = torch.cat([f.reshape(-1) for f in Fstar]).to(device)
        _, y = eng.forward_mn(X, theta_star, w, 0, 50.0, 1.0)
        B0 = MTR.make_BcpInit(list(dims) + [C], R, nn, scale=0.2, device='cpu')
        model = None
        cw = torch.ones(C, device=device)
    n_total = sharder.total(n_local, device)
    eng.set_option('fused', args.fused)
    eng.set_option('flow', args.flow)
    if args.flow_debug:
        eng.set_option('flow_debug', args.flow_debug)
    if args.flow_window_mb:
        eng.set_option('flow_window_mb', args.flow_window_mb)

    theta = (model.theta if model is not None else
             torch.cat([b.reshape(-1) for b in B0]).to(device=device, dtype=torch.float32).contiguous())
    m_, v_, vm_ = torch.zeros_like(theta), torch.zeros_like(theta), torch.zeros_like(theta)
    gs = torch.empty(eng.n_gradsum, dtype=torch.float64, device=device)
    grad = torch.empty_like(theta)
    loss = torch.empty(2, dtype=torch.float64, device=device)
    W_total = n_total                                     # class weights are ones in the bench

    step_no = [0]

    def step():
        """One fit iteration: fwd + grad (2 passes over X) + all-reduce + penalty/normalise + Adam."""
        step_no[0] += 1
        if kind == 'std':
            eng.fwd_grad_std(X, y, theta, w, 0, 50.0, 1.0, gradsum=gs)
            sharder.sum_(gs)
            eng.finish(gs, 2.0 / n_total, 1.0 / n_total, theta, LAMBDA, 0, 50.0, 1.0, grad=grad, loss=loss)
        else:
            eng.fwd_grad_mn(X, y, cw, theta, w, 0, 50.0, 1.0, gradsum=gs)
            sharder.sum_(gs)
            eng.finish(gs, 1.0 / W_total, 1.0 / W_total, theta, LAMBDA, 0, 50.0, 1.0, grad=grad, loss=loss)
        eng.adam_step(theta, grad, m_, v_, vm_, step_no[0], lr=ADAM['lr'])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    eng.profile(True)
    launches0 = eng.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1)
    prof = eng.profile_read()
    eng.profile(False)
    launches = eng.launches - launches0
    launch_info = eng.launch_info()
    n_gradsum = eng.n_gradsum
    final_loss = loss.cpu().tolist()
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (CUDA events on the launching stream, this rank) ----
    peak, peak_src = peaks()
    fwd_ms = prof['fwd_ms'] / max(1, prof['fwd_launches'])
    grad_ms = prof['grad_ms'] / max(1, prof['grad_launches'])
    fused_ms = prof['fused_ms'] / max(1, prof['fused_launches'])
    ratios = {}
    try:
        ratios = json.load(open(os.path.join(ROOT, 'profiles', 'traffic_ratios.json')))
    except Exception:
        pass
    if prof['fused_launches'] > 0:
        # single-pass kernel: does the work of both passes (algorithmic bytes = 2 x bytes(X))
        # while reading X from HBM once -> "achieved" exceeds the HBM peak by design; traffic shows the real bytes
        is_flow = launch_info['path'].startswith('single-launch dataflow')
        dom, dom_ms, alg = ('k_flow' if is_flow else 'k_fused_std'), fused_ms, 2 * x_bytes
        extra = {'k_single_ms': fused_ms, 'hbm_gbs_if_x_read_once': x_bytes / (fused_ms * 1e-3) / 1e9,
                 'hbm_frac_if_x_read_once': x_bytes / (fused_ms * 1e-3) / 1e9 / peak,
                 'note': ('single-launch dataflow kernel: the gradient warps re-read each sample from L2 a bounded '
                          'number of samples behind the forward warps' if is_flow else
                          'single-pass kernel: the second pass over X is served from shared memory') +
                         ', so DRAM traffic approaches 1 x bytes(X) while the algorithmic (2-pass) byte count is '
                         '2 x bytes(X); see traffic for the measured DRAM bytes',
                 'share_of_step': {dom: fused_ms / ms_per_step}}
    else:
        dom = 'k_grad' if grad_ms >= fwd_ms else 'k_fwd'
        dom_ms, alg = max(grad_ms, fwd_ms), x_bytes
        extra = {'k_fwd_ms': fwd_ms, 'k_grad_ms': grad_ms,
                 'k_fwd_gbs': x_bytes / (fwd_ms * 1e-3) / 1e9 if fwd_ms else None,
                 'k_grad_gbs': x_bytes / (grad_ms * 1e-3) / 1e9 if grad_ms else None,
                 'share_of_step': {'k_fwd': fwd_ms / ms_per_step, 'k_grad': grad_ms / ms_per_step}}
    achieved = alg / (dom_ms * 1e-3) / 1e9
    tr = ratios.get(f'{dom}_{kind}') or ratios.get(dom)
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak,
                'traffic': (tr['dram_bytes_per_algorithmic_byte'] * alg) if tr else None,
                'traffic_source': tr['source'] if tr else None,
                'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg,
                'iteration_gbs': 2 * x_bytes / (ms_per_step * 1e-3) / 1e9,
                'iteration_frac_of_2pass_roofline': 2 * x_bytes / (ms_per_step * 1e-3) / 1e9 / peak}
    roofline.update(extra)

    # ---- end to end through the public API with HOST buffers (std workloads) -----------------
    e2e = None
    if not args.no_e2e and kind == 'std':
        pool_n = int(min(n_local, max(256, (2 << 30) // (D * elt))))
        pool = torch.empty((pool_n, *dims), dtype=dt, pin_memory=True)
        pool.copy_(X[:pool_n])
        Xh = CycledHostArray(pool, n_local)
        reps = (n_local + pool_n - 1) // pool_n
        yh = torch.cat([y[:pool_n].cpu()] * reps)[:n_local].contiguous()
        m2 = STR.CP_linear_regression((n_local, *dims), dtype=dt, rank=R, non_negative=False,
                                      Bcp_init=[b.clone() for b in B0], device=device,
                                      shard_group='world' if world > 1 else None)
        chunk = int(max(1, min(pool_n, (1 << 30) // (D * elt))))
        m2.fit_Adam(Xh, yh, lambda_L2=LAMBDA, max_iter=1, tol=0.0, patience=10 ** 9, Adam_kwargs=ADAM,
                    out_of_core=True, chunk_samples=chunk)                       # warm-up pass
        barrier()
        t0 = time.perf_counter()
        m2.fit_Adam(Xh, yh, lambda_L2=LAMBDA, max_iter=args.e2e_steps, tol=0.0, patience=10 ** 9, Adam_kwargs=ADAM,
                    out_of_core=True, chunk_samples=chunk)
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        e2e = {'value': n_total / e2e_s, 'unit': 'samples/s', 'h2d_bytes_per_step': int(x_bytes),
               'd2h_bytes_per_step': 8, 'steps': args.e2e_steps, 'ms_per_step': e2e_s * 1e3,
               'api': 'CP_linear_regression.fit_Adam(X_host, y, out_of_core=True): every iteration streams X '
                      'from pinned host memory (PCIe-bound); host y uploaded once per call (N x 4 B)',
               'h2d_gbs_per_gpu': x_bytes / e2e_s / 1e9,
               'host_buffer': f'{pool_n}-sample pinned pool cycled to N={n_local} (synthetic data)'}
        del pool, m2
    elif not args.no_e2e:
        pool_n = int(min(n_local, max(256, (2 << 30) // (D * elt))))
        pool = torch.empty((pool_n, *dims), dtype=dt, pin_memory=True)
        pool.copy_(X[:pool_n])
        Xh = CycledHostArray(pool, n_local)
        reps = (n_local + pool_n - 1) // pool_n
        yh = torch.cat([y[:pool_n].cpu()] * reps)[:n_local].contiguous()
        chunk = int(max(1, min(pool_n, (1 << 30) // (D * elt))))
        m2 = MTR.CP_logistic_regression(Xh, yh, rank=R, non_negative=False, Bcp_init=[b.clone() for b in B0],
                                        device=device, shard_group='world' if world > 1 else None, n_classes=C,
                                        out_of_core=True, chunk_samples=chunk)
        cwh = np.ones(C, dtype=np.float32)
        m2.fit_Adam(lambda_L2=LAMBDA, max_iter=1, tol=0.0, patience=10 ** 9, weights=cwh, Adam_kwargs=ADAM)   # warm-up
        barrier()
        t0 = time.perf_counter()
        m2.fit_Adam(lambda_L2=LAMBDA, max_iter=args.e2e_steps, tol=0.0, patience=10 ** 9, weights=cwh, Adam_kwargs=ADAM)
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_s = float(te.item())
        e2e = {'value': n_total / e2e_s, 'unit': 'samples/s', 'h2d_bytes_per_step': int(x_bytes),
               'd2h_bytes_per_step': 8, 'steps': args.e2e_steps, 'ms_per_step': e2e_s * 1e3,
               'api': 'CP_logistic_regression(X_host, y, out_of_core=True).fit_Adam(...): every iteration streams X '
                      'from pinned host memory (PCIe-bound); host y uploaded once (N x 8 B)',
               'h2d_gbs_per_gpu': x_bytes / e2e_s / 1e9,
               'host_buffer': f'{pool_n}-sample pinned pool cycled to N={n_local} (synthetic data)'}
        del pool, m2

    # ---- the call a user makes: one fit_Adam(X_host, ...) of K iterations, X uploaded ONCE inside the
    # timed region (pinned, double-buffered) and then resident.  The original X is freed first.
    if e2e is not None and e2e.get('value') is not None:
        pool_n = int(min(n_local, max(256, (2 << 30) // (D * elt))))
        pool = torch.empty((pool_n, *dims), dtype=dt, pin_memory=True)
        pool.copy_(X[:pool_n])
        reps = (n_local + pool_n - 1) // pool_n
        yh = torch.cat([y[:pool_n].cpu()] * reps)[:n_local].contiguous()
        Xh = CycledHostArray(pool, n_local)
        X = None
        model = None
        eng = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        K = max(1, args.steps)
        barrier()
        t0 = time.perf_counter()
        if kind == 'std':
            m3 = STR.CP_linear_regression((n_local, *dims), dtype=dt, rank=R, non_negative=False,
                                          Bcp_init=[b.clone() for b in B0], device=device,
                                          shard_group='world' if world > 1 else None)
            m3.fit_Adam(Xh, yh, lambda_L2=LAMBDA, max_iter=K, tol=0.0, patience=10 ** 9, Adam_kwargs=ADAM)
        else:
            m3 = MTR.CP_logistic_regression(Xh, yh, rank=R, non_negative=False, Bcp_init=[b.clone() for b in B0],
                                            device=device, shard_group='world' if world > 1 else None, n_classes=C)
            m3.fit_Adam(lambda_L2=LAMBDA, max_iter=K, tol=0.0, patience=10 ** 9, weights=np.ones(C, dtype=np.float32),
                        Adam_kwargs=ADAM)
        barrier()
        call_s = time.perf_counter() - t0
        tc = torch.tensor([call_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        call_s = float(tc.item())
        # e2e.value = the reference-facing call with HOST buffers (the reference's own fit_Adam signature, nothing
        # resident beforehand, every host<->device copy inside the timed region).  The variant that re-streams X
        # on every iteration (our out_of_core extension, for X larger than HBM) is kept beside it.
        streaming = e2e
        y_bytes = n_local * (elt if kind == 'std' else 8)
        e2e = {'value': n_total * K / call_s, 'unit': 'samples/s',
               'h2d_bytes_per_step': int((x_bytes + y_bytes) // K), 'd2h_bytes_per_step': 8,
               'iterations': K, 'seconds': call_s, 'ms_per_step': call_s * 1e3 / K,
               'h2d_bytes_total': int(x_bytes + y_bytes), 'final_loss': float(m3.loss_running[-1]),
               'api': ('CP_linear_regression(...).fit_Adam(X_host, y_host, max_iter=K)' if kind == 'std' else
                       'CP_logistic_regression(X_host, y_host, ...).fit_Adam(max_iter=K)') +
                      ': ONE call with the reference signature and K = --steps iterations, timed from construction to '
                      'return: one pinned double-buffered upload of X and y (h2d_bytes_total; h2d_bytes_per_step is '
                      'that total / K), K resident fit iterations, one 8-byte loss read per iteration',
               'host_buffer': streaming['host_buffer'],
               'streaming_every_iteration': streaming}
        del pool, m3

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(wl)

    if rank == 0:
        out = {'metric': 'samples/sec per fit iteration (fwd+grad+step)', 'value': value, 'unit': 'samples/s',
               'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
               'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
               'dtype': 'f32' if dt == torch.float32 else 'f64', 'data': 'synthetic',
               'config': {'workload': DESCR[wl], 'n_per_gpu': n_local, 'n_total': int(n_total),
                          'x_bytes_per_gpu': int(x_bytes), 'optimizer': 'Adam lr=0.01 amsgrad', 'lambda_L2': LAMBDA,
                          'l2_flush': 'not needed: X per GPU (%.1f GB) >> 126 MB L2' % (x_bytes / 1e9)
                          if x_bytes > (1 << 30) else 'X smaller than L2+: numbers are cache-assisted',
                          'parallelism': f'sample-sharded x{world}, one all-reduce of {n_gradsum} doubles/iter',
                          'host_cpus_bound_to_gpu_numa_node': numa,
                          'launch': launch_info},
               'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(launches),
               'clocks': clocks, 'final_loss': final_loss}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
