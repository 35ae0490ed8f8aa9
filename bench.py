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
    # the spectral variant (spectral_tensor_regression.py); dims = (W, D), R = (rank_normal, rank_spectral,
    # complex columns), C = outputs.  Not a BASELINE config: measured as a secondary record.
    'spec1': ('spec', 200000, (64, 128), (2, 2, 2), 4, torch.float32),
    # diagnostic only (not a BASELINE config): the standard model on the cfg3 sample shape
    'dbg3': ('std', 62500, (100, 50, 20), 6, 0, torch.float32),
}
DESCR = {
    'cfg1': 'standard CP regression, X (N=2000, 20,30,40), rank 5, fp32',
    'cfg2': 'standard CP regression, X (N=200000, 64,64,32), rank 8, fp32',
    'cfg3': 'multinomial CP regression, X (N=62500 per GPU, 100,50,20), n_classes=10, rank 6, fp32',
    'cfg4': '5-mode standard CP regression, X (N=100000, 16,16,16,32), rank 12, fp64',
    'cfg5': 'multinomial CP regression, X (N=312500 per GPU, 100,50,20), n_classes=4, rank 4, fp32',
    'spec1': 'spectral CP regression (spectral_tensor_regression.py), X (T=200000 per GPU, 64,128), 4 outputs, rank_normal 2, rank_spectral 2, 2 complex columns, fp32',
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
    if kind == 'spec':
        rn, rs, cc = R
        shapes = [(dims[0], rn), (dims[1], rn), (C, rn), (dims[0], rs * cc), (dims[1], rs), (C, rs)]
        Fstar = [0.1 * torch.randn(shp, generator=gc, dtype=dt) for shp in shapes]
    else:
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
    """--impl reference: the reference's own CPU implementation of the fit_Adam iteration on the host cores, rank 0
    only.  The UNMODIFIED reference modules are used when they can be imported (oracle.ref_loader: $TR_REFERENCE_DIR
    or /root/reference, then baseline/_ref — build container only, kind "reference"); on the GPU box they do not
    exist and the pinned oracle port of the same algorithm runs instead (kind "port")."""
    if rank != 0:
        return
    wl = args.workload
    kind, _, dims, R, C, dt = WORKLOADS[wl]
    from oracle import tr_oracle as O
    from oracle import ref_loader
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D = int(np.prod(dims))
    n_s = int(max(64, min(2000, (1 << 30) // (D * (4 if dt == torch.float32 else 8)))))
    ref_kind, ref_src = 'port', 'oracle port (oracle/tr_oracle.py)'
    for cand in (ref_loader.REFERENCE_DIR, os.path.join(ROOT, 'baseline', '_ref')):
        if os.path.isfile(os.path.join(cand, 'standard_tensor_regression.py')):
            ref_loader.REFERENCE_DIR = cand
            ref_kind, ref_src = 'reference', f'unmodified reference modules from {cand} (tensorly stand-in)'
            break
    print(f'reference arm: using the {ref_src}', file=sys.stderr)
    nn = [False] * (len(dims) + 1)
    if kind == 'std':
        X, y, _ = O.synth_std(n_s, dims, R, 1234 + 2, dtype=dt)
        B0 = O.init_std(dims, R, nn, dtype=dt)
    else:
        X, y, _ = O.synth_mn(n_s, dims, R, C, 1234 + 3)
        B0 = O.init_mn(list(dims) + [C], R, nn)
    if ref_kind == 'reference':
        if kind == 'std':
            ref = ref_loader.standard()
            mdl = ref.CP_linear_regression(X.shape, dtype=dt, rank=R, non_negative=False,
                                           Bcp_init=[b.clone().requires_grad_(True) for b in B0], device='cpu')
            run = lambda k: mdl.fit_Adam(X, y, lambda_L2=LAMBDA, max_iter=k, tol=0.0, patience=10 ** 9,  # noqa: E731
                                         Adam_kwargs=dict(ADAM))
        else:
            ref = ref_loader.multinomial()
            mdl = ref.CP_logistic_regression(X, y, rank=R, non_negative=False,
                                             Bcp_init=[b.clone().requires_grad_(True) for b in B0], device='cpu')
            run = lambda k: mdl.fit_Adam(lambda_L2=LAMBDA, max_iter=k, tol=0.0, patience=10 ** 9,  # noqa: E731
                                         weights=np.ones(C, dtype=np.float32), Adam_kwargs=dict(ADAM))
        if args.warmup > 0:
            run(args.warmup)
        t0 = time.perf_counter()
        run(args.steps)
        dt_s = time.perf_counter() - t0
    else:
        if kind == 'std':
            B = [b.clone().requires_grad_(True) for b in B0]
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
            B = [b.clone().requires_grad_(True) for b in B0]
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
              f'{ref_src}: fit_Adam iteration on torch CPU, {torch.get_num_threads()} threads')
    out = {'impl': 'reference', 'metric': 'samples/sec per fit iteration (fwd+grad+step)', 'value': val,
           'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
           'ms_per_step': dt_s / args.steps * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': 'f32' if dt == torch.float32 else 'f64', 'data': 'synthetic',
           'config': {'workload': DESCR[wl], 'sample': sample},
           'cpu_baseline': {'value': val, 'unit': 'samples/s', 'cores': cores, 'kind': ref_kind, 'sample': sample},
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


def fmt_dtype(dt):
    return 'f32' if dt == torch.float32 else 'f64'


class Workload:
    """One BASELINE config resident on this rank's GPU: data, engine, optimizer state and the fit-iteration step."""

    def __init__(self, wl, n_local, rank, world, device, E, STR, MTR, fused=-1, flow=-1):
        self.wl, self.n_local, self.world, self.device = wl, int(n_local), world, device
        kind, _, dims, R, C, dt = WORKLOADS[wl]
        self.kind, self.dims, self.R, self.C, self.dt = kind, dims, R, C, dt
        self.D = int(np.prod(dims))
        self.elt = 4 if dt == torch.float32 else 8
        self.x_bytes = self.n_local * self.D * self.elt
        self.X, Fstar = make_device_data(wl, self.n_local, rank, device)
        self.sharder = None
        nn = [False] * (len(dims) + 1)
        torch.manual_seed(321)
        if kind == 'spec':
            rn, rs, cc = R
            self.eng = E.SpectralEngine(dims[0], dims[1], C, rn, rs, cc, dt, device)
            self.w = torch.ones(rn + rs, dtype=dt, device=device)
            theta_star = torch.cat([f.reshape(-1) for f in Fstar] + [torch.zeros(C, dtype=dt)]).to(device)
            self.y = self.eng.forward(self.X, theta_star, self.w, 0, 50.0, 1.0, want=('yhat',))['yhat']
            self.y += 0.01 * torch.randn(self.y.shape, dtype=dt, device=device)
            gi = torch.Generator().manual_seed(321)
            self.theta = (0.2 * torch.rand(self.eng.P, generator=gi, dtype=dt) - 0.05).to(device).contiguous()
            self.theta[-C:] = 0
            self.cw = None
        elif kind == 'std':
            self.eng = E.Engine(dims, R, 0, dt, device)
            self.w = torch.ones(R, dtype=dt, device=device)
            theta_star = torch.cat([f.reshape(-1) for f in Fstar] + [torch.tensor([0.1], dtype=dt)]).to(device)
            self.y = self.eng.forward_std(self.X, theta_star, self.w, 0, 50.0, 1.0)
            self.y += 0.01 * torch.randn(self.y.shape, dtype=dt, device=device)
            self.B0 = STR.make_BcpInit(list(dims), R, nn, scale=1, device='cpu', dtype=dt)
            self.theta = torch.cat([b.reshape(-1) for b in self.B0] + [torch.zeros(1, dtype=dt)]).to(device).contiguous()
            self.cw = None
        else:
            self.eng = E.Engine(dims, R, C, torch.float32, device)
            self.w = torch.ones(R, device=device)
            theta_star = torch.cat([f.reshape(-1) for f in Fstar]).to(device)
            _, self.y = self.eng.forward_mn(self.X, theta_star, self.w, 0, 50.0, 1.0)
            self.B0 = MTR.make_BcpInit(list(dims) + [C], R, nn, scale=0.2, device='cpu')
            self.theta = torch.cat([b.reshape(-1) for b in self.B0]).to(device=device, dtype=torch.float32).contiguous()
            self.cw = torch.ones(C, device=device)
        self.sharder = E.ShardedSum(engine=self.eng) if world > 1 else E.ShardedSum(enabled=False)
        if kind != 'spec':
            self.eng.set_option('fused', fused)
            if flow == 1:
                self.eng.set_option('flow', 1)
        self.reset()

    def reset(self):
        th = self.theta
        self.th = th.clone()
        self.m_, self.v_, self.vm_ = torch.zeros_like(th), torch.zeros_like(th), torch.zeros_like(th)
        self.gs = torch.empty(self.eng.n_gradsum, dtype=torch.float64, device=self.device)
        self.grad = torch.empty_like(th)
        self.loss = torch.empty(2, dtype=torch.float64, device=self.device)
        self.step_no = 0

    def step(self, X=None, y=None, n_total=None):
        """One fit iteration: fwd + grad over X + all-reduce + penalty/normalise + Adam."""
        X = self.X if X is None else X
        y = self.y if y is None else y
        self.step_no += 1
        eng = self.eng
        if self.kind == 'spec':
            eng.fwd_grad(X, y, self.th, self.w, 0, 50.0, 1.0, gradsum=self.gs)
            self.sharder.sum_(self.gs)
            nt = n_total * self.C                      # MSE over (T, n_out)
            eng.finish(self.gs, 2.0 / nt, 1.0 / nt, self.th, LAMBDA, 0, 50.0, 1.0, grad=self.grad, loss=self.loss)
        elif self.kind == 'std':
            eng.fwd_grad_std(X, y, self.th, self.w, 0, 50.0, 1.0, gradsum=self.gs)
            self.sharder.sum_(self.gs)
            eng.finish(self.gs, 2.0 / n_total, 1.0 / n_total, self.th, LAMBDA, 0, 50.0, 1.0, grad=self.grad, loss=self.loss)
        else:
            eng.fwd_grad_mn(X, y, self.cw, self.th, self.w, 0, 50.0, 1.0, gradsum=self.gs)
            self.sharder.sum_(self.gs)
            eng.finish(self.gs, 1.0 / n_total, 1.0 / n_total, self.th, LAMBDA, 0, 50.0, 1.0, grad=self.grad, loss=self.loss)
        eng.adam_step(self.th, self.grad, self.m_, self.v_, self.vm_, self.step_no, lr=ADAM['lr'])

    def free(self):
        self.X = None
        self.y = None
        self.eng.close()


def timed_steps(wk, steps, warmup, dist, rank, local_rank, sample_clocks, X=None, y=None, n_total=None):
    """W untimed + K timed steps between barrier + synchronize, CUDA events, max over ranks.  Returns a dict
    with ms_per_step, value, the library's per-kernel event times, launch count / info and the clocks."""
    world = wk.world
    device = wk.device

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if n_total is None:
        n_total = wk.sharder.total(wk.n_local if X is None else X.shape[0], device)
    wk.reset()
    for _ in range(warmup):
        wk.step(X, y, n_total)
    barrier()
    wk.eng.profile(True)
    launches0 = wk.eng.launches
    sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(steps):
        wk.step(X, y, n_total)
    ev1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    prof = wk.eng.profile_read()
    wk.eng.profile(False)
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / steps
    return {'ms_per_step': ms_per_step, 'value': n_total / (ms_per_step * 1e-3), 'n_total': int(n_total), 'prof': prof,
            'launches': wk.eng.launches - launches0, 'launch_info': wk.eng.launch_info(), 'clocks': clocks,
            'final_loss': wk.loss.cpu().tolist(), 'allreduce_via': wk.sharder.via}


def roofline_of(res, wk, x_bytes, ratios):
    """roofline object of the dominant streaming kernel from the library's CUDA-event times (this rank)."""
    peak, peak_src = peaks()
    prof, ms_per_step = res['prof'], res['ms_per_step']
    fwd_ms = prof['fwd_ms'] / max(1, prof['fwd_launches'])
    grad_ms = prof['grad_ms'] / max(1, prof['grad_launches'])
    fused_ms = prof['fused_ms'] / max(1, prof['fused_launches'])
    if prof['fused_launches'] > 0:
        # single-pass kernel: does the work of both passes (algorithmic bytes = 2 x bytes(X)) while
        # reading X from HBM once -> "achieved" exceeds the HBM peak by design; frac_physical and traffic show the
        # bytes that really cross the HBM interface
        path = res['launch_info']['path']
        dom = 'k_flow' if path.startswith('single-launch dataflow') else (
            'k_fused_mn' if wk.kind == 'mn' else ('k_spec_single' if wk.kind == 'spec' else 'k_fused_std'))
        dom_ms, alg, phys = fused_ms, 2 * x_bytes, x_bytes
        extra = {'k_single_ms': fused_ms,
                 'note': 'single-pass kernel: X is read from HBM once, the second (gradient) pass is served from '
                         'cluster shared memory (spectral variant: from the SM\'s own shared memory); frac counts the 2-pass algorithmic bytes (the contract of SURVEY 8d), '
                         'frac_physical the bytes that cross the HBM interface',
                 'share_of_step': {dom: fused_ms / ms_per_step}}
    else:
        dom = 'k_grad' if grad_ms >= fwd_ms else 'k_fwd'
        if wk.kind == 'spec':
            first = 'k_spec_fused' if res['launch_info'].get('df1_slabs') == 0 else 'k_spec_fwd'
            dom = 'k_spec_grad' if grad_ms >= fwd_ms else first
        dom_ms, alg, phys = max(grad_ms, fwd_ms), x_bytes, x_bytes
        extra = {'k_fwd_ms': fwd_ms, 'k_grad_ms': grad_ms,
                 'k_fwd_gbs': x_bytes / (fwd_ms * 1e-3) / 1e9 if fwd_ms else None,
                 'k_grad_gbs': x_bytes / (grad_ms * 1e-3) / 1e9 if grad_ms else None,
                 'share_of_step': {'k_fwd': fwd_ms / ms_per_step, 'k_grad': grad_ms / ms_per_step}}
    achieved = alg / (dom_ms * 1e-3) / 1e9 if dom_ms else None
    tr = ratios.get(f'{dom}_{wk.kind}') or ratios.get(dom)
    out = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
           'frac': achieved / peak if achieved else None,
           'frac_physical': phys / (dom_ms * 1e-3) / 1e9 / peak if dom_ms else None,
           'physical_gbs': phys / (dom_ms * 1e-3) / 1e9 if dom_ms else None,
           'traffic': (tr['dram_bytes_per_algorithmic_byte'] * alg) if tr else None,
           'traffic_source': tr['source'] if tr else None,
           'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg,
           'iteration_gbs': 2 * x_bytes / (ms_per_step * 1e-3) / 1e9,
           'iteration_frac_of_2pass_roofline': 2 * x_bytes / (ms_per_step * 1e-3) / 1e9 / peak}
    out.update(extra)
    return out


def multi_gpu_check(rank, world, device, dist, E, STR, MTR):
    """Do N NCCL ranks compute the single-GPU answer?  Every rank builds the SAME small global dataset from a CPU
    seed; rank g fits its shard_bounds slice for K Adam steps with shard_group='world' (one all-reduce of the packed
    gradient sums per step); rank 0 also fits the whole set alone.  Reports the norm-relative difference of the
    fitted parameters / losses and whether the replicas' parameters are bit-identical after the K steps."""
    from oracle import tr_oracle as O            # synthetic data generators only (checker side)
    K = 5
    out = {'steps': K, 'world': world}
    ok = True
    cases = [('std_f32_single_pass', 'std', (64, 64, 32), 8, 0, torch.float32, 1, 1e-6),
             ('std_f64_two_pass', 'std', (16, 16, 16, 32), 12, 0, torch.float64, 0, 1e-12),
             ('mn_f32', 'mn', (100, 50, 20), 6, 10, torch.float32, -1, 1e-6),
             # spectral variant: dims = (W, D), R = (rank_normal, rank_spectral, complex columns), C = outputs
             ('spec_f32', 'spec', (16, 128), (2, 2, 2), 3, torch.float32, -1, 1e-6),
             # ... the same on the single-pass kernel (k_spec_single; fused = 2 forces it at this size)
             ('spec_f32_single_pass', 'spec', (16, 128), (2, 2, 2), 3, torch.float32, 2, 1e-6),
             ('spec_f64', 'spec', (12, 40), (1, 2, 3), 2, torch.float64, -1, 1e-12)]
    for name, kind, dims, R, C, dt, fused, tol in cases:
        n_glob = 24 * world + 5                         # uneven split on purpose
        nn = [False] * (len(dims) + 1)
        if kind == 'spec':
            from oracle import tr_oracle_spectral as OS
            from tensor_regression_b200 import spectral_tensor_regression as SPR
            rn, rs, cc = R
            X, y = OS.synth(n_glob, dims[0], dims[1], C, rn, rs, cc, 4244, dtype=dt)
            B0 = OS.init(dims[0], dims[1], C, rn, rs, cc, dtype=dt, seed=11)
        elif kind == 'std':
            X, y, _ = O.synth_std(n_glob, dims, R, 4242, dtype=dt)
            B0 = O.init_std(dims, R, nn, dtype=dt)
        else:
            X, y, _ = O.synth_mn(n_glob, dims, R, C, 4243)
            B0 = O.init_mn(list(dims) + [C], R, nn, scale=0.2)
        lo, hi = E.shard_bounds(n_glob, rank, world)

        def fit(Xs, ys, group):
            if kind == 'spec':
                m = SPR.CP_linear_regression((Xs.shape[0], *dims), (Xs.shape[0], C), dtype=dt, rank_normal=R[0],
                                             rank_spectral=R[1], n_complex_dim=R[2] - 1,
                                             Bcp_init=[[b.clone() for b in B0[0]], [b.clone() for b in B0[1]]],
                                             device=device, shard_group=group)
                if fused == 2:
                    m._engine().set_option('spec_single', 1)
                m.fit_Adam(Xs.to(device), ys.to(device), lambda_L2=LAMBDA, max_iter=K, tol=0.0, patience=10 ** 9,
                           Adam_kwargs=ADAM)
                if fused == 2:
                    assert m._engine().launch_info()['path'].startswith('single-pass')
            elif kind == 'std':
                m = STR.CP_linear_regression((Xs.shape[0], *dims), dtype=dt, rank=R, Bcp_init=[b.clone() for b in B0],
                                             device=device, shard_group=group)
                if fused >= 0:
                    m._engine().set_option('fused', fused)
                m.fit_Adam(Xs.to(device), ys.to(device), lambda_L2=LAMBDA, max_iter=K, tol=0.0, patience=10 ** 9,
                           Adam_kwargs=ADAM)
            else:
                m = MTR.CP_logistic_regression(Xs.to(device), ys, rank=R, Bcp_init=[b.clone() for b in B0], device=device,
                                               shard_group=group, n_classes=C)
                m.fit_Adam(lambda_L2=LAMBDA, max_iter=K, tol=0.0, patience=10 ** 9,
                           weights=np.ones(C, dtype=np.float32), Adam_kwargs=ADAM)
            th, losses = m.theta.detach().clone(), list(m.loss_running)
            if kind == 'mn':
                m.X = None
            m.close()
            return th, losses

        th_sh, loss_sh = fit(X[lo:hi], y[lo:hi], 'world')
        gathered = [torch.empty_like(th_sh) for _ in range(world)]
        dist.all_gather(gathered, th_sh)
        identical = all(torch.equal(gathered[0], g) for g in gathered)
        rec = {'n_global': n_glob, 'replicas_bit_identical': bool(identical), 'tol': tol}
        if rank == 0:
            th_1, loss_1 = fit(X, y, None)
            rec['theta_rel_diff'] = float((th_sh - th_1).abs().max() / th_1.abs().max())
            rec['loss_rel_diff'] = float(max(abs(a - b) for a, b in zip(loss_sh, loss_1)) / abs(loss_1[0]))
            rec['pass'] = bool(identical and rec['theta_rel_diff'] <= tol * 10 and rec['loss_rel_diff'] <= tol)
            ok = ok and rec['pass']
        dist.barrier()
        out[name] = rec
    out['pass'] = bool(ok)
    return out


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
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help="weak: every rank holds the workload's per-GPU N (default); strong: the workload's N is split over the ranks")
    ap.add_argument('--secondary', default='cfg3,cfg4,cfg5,spec1', help='comma list of further workloads measured after the primary (device-timed only); "" = none')
    ap.add_argument('--n-local', type=int, default=0, help='override samples per GPU (debug)')
    ap.add_argument('--secondary-n-local', type=int, default=0, help='override samples per GPU of the secondary workloads (debug)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-check', action='store_true', help='skip the multi-rank parity check')
    ap.add_argument('--e2e-steps', type=int, default=2)
    ap.add_argument('--pageable-gib', type=float, default=8.0, help='size of the plain numpy array of the pageable e2e leg')
    ap.add_argument('--fused', type=int, default=-1, help='-1 auto, 0 never, 1 always: single-pass cluster kernels')
    ap.add_argument('--flow', type=int, default=-1, help='1: experimental dataflow kernel (needs a FLOW=1 build)')
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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxrank(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ratios = {}
    try:
        ratios = json.load(open(os.path.join(ROOT, 'profiles', 'traffic_ratios.json')))
    except Exception:
        pass

    # ---- multi-rank parity: N NCCL ranks == one GPU, replicas bit-identical ----------------------
    check = None
    if world > 1 and not args.no_check:
        check = multi_gpu_check(rank, world, device, dist, E, STR, MTR)

    wl = args.workload
    kind, n_cfg, dims, R, C, dt = WORKLOADS[wl]
    n_local = args.n_local or n_cfg
    if args.scaling == 'strong':
        lo, hi = E.shard_bounds(n_local, rank, world)
        n_local = hi - lo

    # ---- primary workload, X resident in HBM ---------------------------------------------------
    wk = Workload(wl, n_local, rank, world, device, E, STR, MTR, fused=args.fused, flow=args.flow)
    D, elt, x_bytes = wk.D, wk.elt, wk.x_bytes
    res = timed_steps(wk, args.steps, args.warmup, dist, rank, local_rank, True)
    n_total, ms_per_step, value = res['n_total'], res['ms_per_step'], res['value']
    roofline = roofline_of(res, wk, x_bytes, ratios)
    n_gradsum = wk.eng.n_gradsum

    # ---- strong scaling beside the weak number: the workload's single-GPU N split over the ranks -------------
    strong = None
    if args.scaling == 'weak' and world > 1:
        lo, hi = E.shard_bounds(n_cfg if not args.n_local else args.n_local, rank, world)
        ns = hi - lo
        rs = timed_steps(wk, args.steps, args.warmup, dist, rank, local_rank, False, X=wk.X[:ns], y=wk.y[:ns])
        strong = {'n_total': rs['n_total'], 'n_per_gpu': ns, 'ms_per_step': rs['ms_per_step'], 'value': rs['value'],
                  'unit': 'samples/s', 'x_bytes_per_gpu': int(ns * D * elt),
                  'roofline': {k: v for k, v in roofline_of(rs, wk, ns * D * elt, ratios).items()
                               if k in ('kernel', 'achieved', 'frac', 'frac_physical', 'k_single_ms', 'k_fwd_ms', 'k_grad_ms',
                                        'share_of_step', 'iteration_frac_of_2pass_roofline')},
                  'note': 'fixed total N (the single-GPU workload) split over the ranks; the part of a step outside the '
                          'streaming kernel(s) = small kernels + launches + the all-reduce (1 - sum(share_of_step))'}
    elif args.scaling == 'weak' and world == 1:
        # one GPU: what a 1/8 shard of the workload costs per iteration (the strong-scaling tail, measured alone)
        ns = max(1, (n_cfg if not args.n_local else args.n_local) // 8)
        rs = timed_steps(wk, args.steps, args.warmup, dist, rank, local_rank, False, X=wk.X[:ns], y=wk.y[:ns])
        strong = {'proxy': 'one GPU running a 1/8 shard of the workload (what each of 8 ranks does under strong scaling, '
                           'without the all-reduce)', 'n_per_gpu': ns, 'ms_per_step': rs['ms_per_step'],
                  'ms_per_step_if_linear': ms_per_step * ns / n_local,
                  'efficiency_upper_bound': (ms_per_step * ns / n_local) / rs['ms_per_step'],
                  'share_of_step': roofline_of(rs, wk, ns * D * elt, ratios)['share_of_step']}

    # ---- end to end through the public API with HOST buffers ------------------------------------------------
    e2e = None
    X, y, B0 = wk.X, wk.y, getattr(wk, 'B0', None)
    if not args.no_e2e:
        pool_n = int(min(n_local, max(256, (2 << 30) // (D * elt))))
        pool = torch.empty((pool_n, *dims), dtype=dt, pin_memory=True)
        pool.copy_(X[:pool_n])
        Xh = CycledHostArray(pool, n_local)
        reps = (n_local + pool_n - 1) // pool_n
        yh = torch.cat([y[:pool_n].cpu()] * reps)[:n_local].contiguous()
        chunk = int(max(1, min(pool_n, (1 << 30) // (D * elt))))
        group = 'world' if world > 1 else None
        cwh = np.ones(max(C, 1), dtype=np.float32)

        def fit_call(Xarg, yarg, K, **kw):
            """the reference-facing call: construct the estimator and run fit_Adam with K iterations"""
            if kind == 'std':
                m = STR.CP_linear_regression((Xarg.shape[0], *dims), dtype=dt, rank=R, non_negative=False,
                                             Bcp_init=[b.clone() for b in B0], device=device, shard_group=group)
                m.fit_Adam(Xarg, yarg, lambda_L2=LAMBDA, max_iter=K, tol=0.0, patience=10 ** 9, Adam_kwargs=ADAM, **kw)
            else:
                m = MTR.CP_logistic_regression(Xarg, yarg, rank=R, non_negative=False, Bcp_init=[b.clone() for b in B0],
                                               device=device, shard_group=group, n_classes=C, **kw)
                m.fit_Adam(lambda_L2=LAMBDA, max_iter=K, tol=0.0, patience=10 ** 9, weights=cwh, Adam_kwargs=ADAM)
            return m

        # (a) X larger than HBM / streamed on EVERY iteration (our out_of_core extension): PCIe-bound
        fit_call(Xh, yh, 1, out_of_core=True, chunk_samples=chunk).close()              # warm-up pass
        barrier()
        t0 = time.perf_counter()
        fit_call(Xh, yh, args.e2e_steps, out_of_core=True, chunk_samples=chunk).close()
        barrier()
        e2e_s = maxrank((time.perf_counter() - t0) / args.e2e_steps)
        streaming = {'value': n_total / e2e_s, 'unit': 'samples/s', 'h2d_bytes_per_step': int(x_bytes),
                     'd2h_bytes_per_step': 8, 'steps': args.e2e_steps, 'ms_per_step': e2e_s * 1e3,
                     'api': 'fit_Adam(X_host, y, out_of_core=True): every iteration streams X from pinned host memory',
                     'h2d_gbs_per_gpu': x_bytes / e2e_s / 1e9}

        # (b) the call a user makes: ONE fit_Adam(X_host, ...) of K iterations, X uploaded once inside the timed
        # region and then resident.  The resident copy of the device-timed legs is freed first.
        wk.free()
        X = y = None
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        K = max(1, args.steps)
        barrier()
        t0 = time.perf_counter()
        m3 = fit_call(Xh, yh, K)
        barrier()
        call_s = maxrank(time.perf_counter() - t0)
        y_bytes = n_local * (elt if kind == 'std' else 8)
        e2e = {'value': n_total * K / call_s, 'unit': 'samples/s',
               'h2d_bytes_per_step': int((x_bytes + y_bytes) // K), 'd2h_bytes_per_step': 8,
               'iterations': K, 'seconds': call_s, 'ms_per_step': call_s * 1e3 / K,
               'h2d_bytes_total': int(x_bytes + y_bytes), 'final_loss': float(m3.loss_running[-1]),
               'h2d_gbs_per_gpu_incl_compute': (x_bytes + y_bytes) / call_s / 1e9,
               'api': ('CP_linear_regression(...).fit_Adam(X_host, y_host, max_iter=K)' if kind == 'std' else
                       'CP_logistic_regression(X_host, y_host, ...).fit_Adam(max_iter=K)') +
                      ': ONE call with the reference signature and K = --steps iterations, timed from construction to '
                      'return: one upload of X and y from pinned host memory (h2d_bytes_total; h2d_bytes_per_step is '
                      'that total / K), K resident fit iterations, one 8-byte loss read per iteration',
               'host_buffer': f'{pool_n}-sample pinned pool cycled to N={n_local} (synthetic data)',
               'streaming_every_iteration': streaming}
        m3.close()
        if kind == 'mn':
            m3.X = None
        del m3
        gc.collect()
        torch.cuda.empty_cache()

        # (c) the same call with what a user of the reference really holds: a plain, PAGEABLE numpy array
        n_pg = int(min(n_local, max(64, int(args.pageable_gib * (1 << 30)) // (D * elt))))
        if n_pg > 0 and args.pageable_gib > 0:
            Xnp = np.empty((n_pg, *dims), dtype=np.float32 if dt == torch.float32 else np.float64)
            pn = pool.numpy()
            for lo_ in range(0, n_pg, pool_n):
                hi_ = min(n_pg, lo_ + pool_n)
                Xnp[lo_:hi_] = pn[:hi_ - lo_]
            ynp = yh[:n_pg].numpy().copy()
            fit_call(Xnp[:max(1, n_pg // 16)], ynp[:max(1, n_pg // 16)], 1).close()     # warm-up (staging ring, threads)
            barrier()
            t0 = time.perf_counter()
            m4 = fit_call(Xnp, ynp, K)
            barrier()
            pg_s = maxrank(time.perf_counter() - t0)
            up = E.upload_stats()
            pg_bytes = n_pg * D * elt
            # the pinned-source figure for the SAME size, for a like-for-like ratio
            m4.close()
            del m4
            gc.collect()
            torch.cuda.empty_cache()
            Xpin = CycledHostArray(pool, n_pg)
            barrier()
            t0 = time.perf_counter()
            m5 = fit_call(Xpin, yh[:n_pg], K)
            barrier()
            pin_s = maxrank(time.perf_counter() - t0)
            m5.close()
            del m5
            e2e['pageable_numpy'] = {
                'value': n_pg * world * K / pg_s, 'unit': 'samples/s', 'n_per_gpu': n_pg, 'iterations': K,
                'seconds': pg_s, 'h2d_bytes_total': int(pg_bytes),
                'upload_seconds': up['seconds'], 'upload_gbs': pg_bytes / up['seconds'] / 1e9 if up['seconds'] else None,
                'host_fill_seconds': up['host_fill_seconds'], 'upload_threads': up['threads'],
                'same_size_from_pinned_pool_seconds': pin_s, 'pageable_over_pinned_time': pg_s / pin_s,
                'api': 'the same fit_Adam call with X a plain np.ndarray (pageable, %.1f GiB): tr_upload stages it through '
                       'a ring of pinned buffers filled by several host threads' % (pg_bytes / 2 ** 30)}
            del Xnp, Xpin
        del pool
    else:
        wk.free()
    gc_collect()

    # ---- further BASELINE configs, device-timed (the north star's multinomial / fp64 / 1 TB configurations) ----
    secondary = []
    for swl in [w_ for w_ in args.secondary.split(',') if w_ and w_ != wl and w_ in WORKLOADS]:
        skind, sn, sdims, sR, sC, sdt = WORKLOADS[swl]
        sn = args.secondary_n_local or sn
        if args.scaling == 'strong':
            lo, hi = E.shard_bounds(sn, rank, world)
            sn = hi - lo
        need = sn * int(np.prod(sdims)) * (4 if sdt == torch.float32 else 8) + (6 << 30)
        free_b, _ = torch.cuda.mem_get_info()
        enough = maxrank(0.0 if free_b >= need else 1.0) == 0.0
        if not enough:
            secondary.append({'workload': DESCR[swl], 'skipped': 'not enough free HBM (%.0f GB needed)' % (need / 1e9)})
            continue
        sw = Workload(swl, sn, rank, world, device, E, STR, MTR, fused=args.fused)
        sr = timed_steps(sw, args.steps, args.warmup, dist, rank, local_rank, False)
        secondary.append({'workload': DESCR[swl], 'name': swl, 'value': sr['value'], 'unit': 'samples/s',
                          'ms_per_step': sr['ms_per_step'], 'n_per_gpu': sn, 'n_total': sr['n_total'],
                          'dtype': fmt_dtype(sdt), 'x_bytes_per_gpu': int(sw.x_bytes), 'scaling': args.scaling,
                          'roofline': roofline_of(sr, sw, sw.x_bytes, ratios), 'gpu_launches': int(sr['launches']),
                          'launch': sr['launch_info'], 'final_loss': sr['final_loss']})
        sw.free()
        del sw
        gc_collect()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(wl)

    if rank == 0:
        out = {'metric': 'samples/sec per fit iteration (fwd+grad+step)', 'value': value, 'unit': 'samples/s',
               'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step,
               'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
               'dtype': fmt_dtype(dt), 'data': 'synthetic',
               'config': {'workload': DESCR[wl], 'n_per_gpu': n_local, 'n_total': int(n_total),
                          'x_bytes_per_gpu': int(x_bytes), 'optimizer': 'Adam lr=0.01 amsgrad', 'lambda_L2': LAMBDA,
                          'l2_flush': 'not needed: X per GPU (%.1f GB) >> 126 MB L2' % (x_bytes / 1e9)
                          if x_bytes > (1 << 30) else 'X smaller than L2+: numbers are cache-assisted',
                          'parallelism': f'sample-sharded x{world}, one all-reduce of {n_gradsum} doubles/iter',
                          'allreduce': res['allreduce_via'],
                          'host_cpus_bound_to_gpu_numa_node': numa,
                          'launch': res['launch_info']},
               'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(res['launches']),
               'clocks': res['clocks'], 'final_loss': res['final_loss'], 'strong': strong,
               'multi_gpu_check': check, 'secondary': secondary}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def gc_collect():
    import gc
    gc.collect()
    torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
