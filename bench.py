#!/usr/bin/env python3
"""bench.py -- batched TSID control updates ("QP solves") per second on N B200s of one node.

  python bench.py --gpus 1 --steps 20 --warmup 3
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W
  python bench.py --impl reference ...      # the CPU path (oracle, one instance per thread) on the same workload

A "step" is one pass of the whole control cycle (kinematics / dynamics, task and contact stacking, the
Goldfarb-Idnani QP, decode, integrate: the k_fast + k_slow launch pair per contact mask, or k_cycle for the
arms) over one batch of synthetic randomized robot states (SURVEY.md 8(d) recipes).  The walking workload
adds the device-side behaviour (references, contact switches, bucketing) to every step.  Weak scaling: every
GPU steps its own batch of the same size, no collective on the data path.

Printed JSON (one line, rank 0):
  value      whole-job updates/s with inputs resident in HBM, CUDA events on the launching stream,
             max over ranks, L2 flushed between timed steps
  e2e        the same metric through the reference-facing call `Batch.update` (C ABI
             iwbc_batch_update): pinned HOST buffers in, HOST results out, copies inside the timed region
  roofline   FP64 FMA roofline of the fused kernel (this path is FP64 CUDA-core bound, not HBM or
             tensor bound: SURVEY.md 8(d)); peak measured live with a dependent-FMA probe; the HBM
             figure for the state/result traffic is given as a secondary number
  cpu_baseline  the CPU oracle (port of the reference path) on this box's host cores, bounded sample
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

OPEN_LOOP = {"stabilizer": {"activated": False}, "collision_detection": {"activated": False}, "closed_loop": False}
WORKLOADS = {
    # name: (controller yaml, overrides, per-GPU batch, recipe)
    # (stabilizer / torque collision detection off and open loop, as the reference's own controller-only test runs it:
    # tests/test_determinism.cpp:66-69)
    "talos_ds_16384": ("talos/talos_pos_tracker.yaml", OPEN_LOOP, 16384, "talos"),
    # config 3, fast-URDF variant (the reference's `-f` runs: nq = 39, QP n = 62, me = 18, mi = 196, 85 cost rows)
    "talos_fast_ds_16384": ("talos/talos_pos_tracker.yaml", dict(OPEN_LOOP, urdf="talos/talos_fast.urdf"), 16384, "talos"),
    "franka_65536": ("franka/pos_tracker.yaml", None, 65536, "franka"),
    # config 5: Talos walking, a mix of double- and single-support shapes bucketed by contact mask
    "talos_walk_131072": ("talos/talos_pos_tracker.yaml", OPEN_LOOP, 131072, "walk"),
    # the same shapes from the CPU-only stand-in recipe (make_walk_masks): what the walking workload was before the device-side behaviour
    "talos_walk_standin_131072": ("talos/talos_pos_tracker.yaml", OPEN_LOOP, 131072, "walk_standin"),
    "icub_ds_16384": ("icub/humanoid_pos_tracker.yaml", OPEN_LOOP, 16384, "talos"),
    "tiago_65536": ("tiago/pos_tracker.yaml", {"urdf": "tiago/tiago_steel_no_wheel.urdf"}, 65536, "arm"),
}
DEFAULT_WORKLOAD = "talos_ds_16384"


# ----------------------------------------------------------------------------- synthetic inputs
def make_inputs(P, B, recipe, seed):
    """config 2 (Franka) / config 3 (Talos double support) randomized states and targets."""
    rng = np.random.default_rng(seed)
    q = np.tile(P.q0(), (B, 1))
    v = np.zeros((B, P.nv))
    refs = np.tile(P.default_refs(), (B, 1))
    na = P.na
    if recipe in ("talos", "walk", "walk_standin"):
        dofs = P.names("dofs")
        mimics = set(P.names("mimics"))
        free = np.array([d not in mimics for d in dofs])
        q[:, P.nq - na:][:, free] += rng.normal(0, 0.02, (B, int(free.sum())))
        v[:, P.nv - na:] = rng.normal(0, 0.05, (B, na))
        off, _ = P.ref_slot("com")
        refs[:, off + 2] += rng.uniform(-0.2, 0.0, B) if recipe == "talos" else 0.0
    else:
        from inria_wbc_b200 import trajs
        S = P.struct()
        lo, hi = np.array(S.q_lo[:na]), np.array(S.q_hi[:na])
        narm = 7 if recipe == "franka" else na
        mg = np.minimum(0.1, 0.25 * (hi[:narm] - lo[:narm]))  # short-range joints (grippers, torso lift) keep a proportional margin
        qa = rng.uniform(lo[:narm] + mg, hi[:narm] - mg, (B, narm))
        q[:, P.nq - na:P.nq - na + narm] = qa
        v[:, P.nv - na:P.nv - na + narm] = rng.normal(0, 0.1, (B, narm))
        # ee reference = construction-time pose composed with a random offset (cheap stand-in for FK(q) o delta)
        off, _ = P.ref_slot("ee")
        p0, R0 = trajs.sample_se3(refs[0, off:off + 24])
        dp = rng.uniform(-0.2, 0.2, (B, 3))
        ang = rng.uniform(0, 0.5, B)
        ax = rng.normal(size=(B, 3))
        ax /= np.linalg.norm(ax, axis=1, keepdims=True)
        for i in range(B):
            R = R0 @ trajs._rot(ax[i], ang[i])
            refs[i, off:off + 24] = trajs.se3_sample(p0 + dp[i], R)
    return q, v, refs


def make_walk_masks(P, B, refs, seed):
    """Config 5 stand-in for a uniformly random tick of humanoid::walk (etc/talos/walk.yaml; src/behaviors/humanoid/walk.cpp:42-60):
    half of the instances in double support, a quarter on each foot; the swing foot's SE3 reference is lifted (step_height
    0.05) and the CoM reference moved over the support foot.  Instances are ordered by mask so that every launch is
    shape-uniform."""
    rng = np.random.default_rng(seed)
    mask = rng.choice(np.array([3, 3, 1, 2], dtype=np.uint32), B)
    mask.sort()
    mask = mask[::-1].copy()  # 3 (double support) first
    lf, _ = P.ref_slot("lf")
    rf, _ = P.ref_slot("rf")
    com, _ = P.ref_slot("com")
    cl, cr = P.contact_index("contact_lfoot"), P.contact_index("contact_rfoot")
    lift = rng.uniform(0.0, 0.05, B)
    for m, swing, support in ((1 << cl, rf, lf), (1 << cr, lf, rf)):
        sel = mask == m
        refs[sel, swing + 2] += lift[sel]
        refs[sel, com:com + 2] = refs[sel, support:support + 2]
    return mask


def make_walk_workload(P, bt, B, seed, device, ncycles=2, stride=20):
    """Config 5: instance i sits at a uniformly random tick of the humanoid::walk cycle (etc/talos/walk.yaml).
    To make (q, v) consistent with that tick's references, ONE instance is first walked through the whole behaviour on the
    device (closed device-side roll-out, open loop as tests/test_determinism.cpp runs the controller) and its state is kept
    every `stride` ticks; instance i = a random snapshot + small noise.  Returns host arrays ordered by contact mask (double
    support first) so that the host-buffer call finds contiguous runs, the per-instance phase, and the behaviour."""
    import torch
    import inria_wbc_b200 as iw
    beh = iw.Behavior(P, "talos/walk.yaml", overrides={"num_of_cycles": ncycles})
    dev = torch.device("cuda", device)
    q1 = torch.from_numpy(P.q0()[None].copy()).to(dev)
    v1 = torch.zeros(1, P.nv, dtype=torch.float64, device=dev)
    r1 = torch.empty(1, P.nref, dtype=torch.float64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    nsnap = beh.total_ticks // stride
    snap_q = torch.empty(nsnap, P.nq, dtype=torch.float64, device=dev)
    snap_v = torch.empty(nsnap, P.nv, dtype=torch.float64, device=dev)
    worst = torch.zeros(1, dtype=torch.int32, device=dev)
    ts = torch.cuda.Stream(device=dev)  # an explicit stream (handle 0 means "the batch's own stream" to the C ABI): copies and ticks stay ordered
    stream = ts.cuda_stream
    torch.cuda.synchronize()
    with torch.cuda.stream(ts):
        for k in range(nsnap):
            snap_q[k].copy_(q1[0])
            snap_v[k].copy_(v1[0])
            bt.behavior_rollout_device(beh, 1, q1.data_ptr(), v1.data_ptr(), r1.data_ptr(), None, k * stride, stride, init_refs=(k == 0),
                                       out_ptrs={"status": st.data_ptr()}, stream=stream)
            worst = torch.maximum(worst, st.abs())
    torch.cuda.synchronize()
    if int(worst[0]) != 0:
        raise SystemExit("bench.py: the seed walk left the OPTIMAL status (%d)" % int(worst[0]))
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, nsnap, B)
    phase = (pick * stride).astype(np.int32)
    # contact mask of every instance at its tick -> order by mask, double support first
    d_phase = torch.from_numpy(phase).to(dev)
    d_refs = torch.empty(B, P.nref, dtype=torch.float64, device=dev)
    d_mask = torch.zeros(B, dtype=torch.int32, device=dev)
    beh.eval_device(device, B, d_phase.data_ptr(), 0, d_refs.data_ptr(), d_mask.data_ptr(), init_refs=True, stream=stream)
    torch.cuda.synchronize()
    masks = d_mask.cpu().numpy().astype(np.uint32)
    order = np.argsort(-masks.astype(np.int64), kind="stable")
    pick, phase, masks = pick[order], phase[order], masks[order]
    refs = d_refs.cpu().numpy()[order]
    q = snap_q.cpu().numpy()[pick]
    v = snap_v.cpu().numpy()[pick]
    mimics = set(P.names("mimics"))
    free = np.array([d not in mimics for d in P.names("dofs")])
    q[:, P.nq - P.na:][:, free] += rng.normal(0, 0.002, (B, int(free.sum())))
    v[:, P.nv - P.na:] += rng.normal(0, 0.01, (B, P.na))
    return np.ascontiguousarray(q), np.ascontiguousarray(v), np.ascontiguousarray(refs), masks, phase, beh


# ----------------------------------------------------------------------------- flop model
def algorithmic_flops(nv, na, n, me, mi, rows, adds, nj, nframes_jac, nsc_pairs):
    """Reference-algorithm FLOPs per update (FMA = 2): SURVEY.md 8(d) F = F_dyn + F_asm + F_qp."""
    f_qp = rows * n * (n + 1) + 2 * rows * n + n ** 3 / 3 + n ** 3 / 3 + 2 * n * n
    for iq in range(me):
        f_qp += 2 * n * n + 2 * n * (n - iq) + iq * iq + 6 * n * (n - iq - 1) + 4 * n
    iq = me
    f_qp += adds * (2 * mi * n + 2 * n * n + 2 * n * (n - iq) + iq * iq + 6 * n * (n - iq - 1) + 6 * n)
    f_qp += 2 * mi * n
    # textbook counts: per joint forward pass (placement, velocity, 2 accelerations, force) ~ 330, backward (RNEA + CRBA
    # force-set propagation over the subtree) ~ 60 per (joint, supported dof) pair ~ nv*depth, Jacobian columns 30/dof,
    # CoM / centroidal 120/joint
    f_dyn = nj * (330 + 120) + nv * 30 + nv * 8 * 60
    # task assembly: one 6 x nv local frame Jacobian per SE3 task / contact (66 flops per column), 3 x nv world Jacobians for
    # the self-collision pairs (2 per pair, 15 flops per column), actuation-bounds mapping C*[M_a | -J_a'] as the reference
    # does it (dense na x na x n), Jc = T' J per contact (12 x 6 x nv)
    f_asm = nframes_jac * nv * 66 + nsc_pairs * 2 * nv * 15 + (2 * na * na * n if mi > 2 * na else 0) + (n - nv) // 12 * (2 * 12 * 6 * nv)
    return dict(f_qp=float(f_qp), f_dyn=float(f_dyn), f_asm=float(f_asm), total=float(f_qp + f_dyn + f_asm))


def problem_counts(P):
    S = P.struct()
    nfr = 0
    nsc = 0
    for k in range(S.ntasks):
        t = S.tasks[k]
        if t.type == 0:
            nfr += 1
        if t.type == 4:
            nsc += t.nav
    nfr += P.ncontacts
    return S.nj, nfr, nsc


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML from a thread every 5 ms; nvidia-smi -lms as the
    fallback when the NVML binding is missing)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = None
        self.nvml = None
        self.run = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.mx = [float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))]
            self.nvml = pynvml
            self.run = True
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while self.run:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                bits = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for nm, bit in self.BITS.items():
                    if bits & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.nvml:
            self.run = False
            self.th.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, 5 ms period, during the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 20"}


# ----------------------------------------------------------------------------- CPU arm
def oracle_rate(P, q, v, refs, seconds, threads):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    n = q.shape[0]
    ol.cycle_batch(P, q[:min(n, 256)], v[:min(n, 256)], refs[:min(n, 256)], nthreads=threads)  # warm-up
    done, t0 = 0, time.perf_counter()
    while True:
        ol.cycle_batch(P, q, v, refs, nthreads=threads)
        done += n
        el = time.perf_counter() - t0
        if el >= seconds:
            break
    return done / el, done, el


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port; the reference itself cannot be built here:
    SURVEY.md 8(c)), one independent controller instance per thread over all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the oracle's own configuration reader (PyYAML + xml.etree): this arm never loads libiwbc_b200.so
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import problem_loader
    yaml, ov, B, recipe = WORKLOADS[args.workload]
    P = problem_loader.load(yaml, overrides=ov)
    threads = os.cpu_count() or 1
    sample = min(B, 2048)
    q, v, refs = make_inputs(P, sample, recipe, 0xB200 + 3)
    masks = None
    if recipe in ("walk", "walk_standin"):  # CPU-only stand-in for the walk's mix of contact shapes (the GPU arm walks a seed instance)
        P.set_task_weight("momentum", 0.0)
        masks = make_walk_masks(P, sample, refs, 0xB200 + 5)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    for _ in range(max(args.warmup, 1)):
        ol.cycle_batch(P, q, v, refs, contact_mask=masks, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ol.cycle_batch(P, q, v, refs, contact_mask=masks, nthreads=threads)
    el = time.perf_counter() - t0
    val = sample * args.steps / el
    line = {"impl": "reference", "metric": "batched TSID control updates (QP solves) per second", "value": val, "unit": "updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "per_step_sample": sample,
                       "note": "CPU oracle port of the reference path (oracle/wbc_oracle.cpp fed by oracle/problem_loader.py); one instance per thread; "
                               "each step is a bounded sample of the workload's recipe"},
            "cpu_baseline": {"value": val, "unit": "updates/s", "cores": threads, "kind": "port", "sample": "%d instances x %d steps" % (sample, args.steps)},
            "e2e": {"value": val, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
MIXED = {"mixed_icub_tiago_16384": ("icub_ds_16384", "tiago_65536", 8192)}  # BASELINE config 4: 50 / 50 iCub + Tiago on one GPU
SECONDARY = ("franka_65536", "tiago_65536", "icub_ds_16384", "mixed_icub_tiago_16384", "talos_walk_131072")


class Part:
    """One robot's share of a workload on this rank's GPU: its handle, synthetic states and device buffers."""

    def __init__(self, iw, torch, workload, B, rank, local, dev):
        yaml, ov, _, recipe = WORKLOADS[workload]
        self.workload, self.yaml, self.recipe, self.B = workload, yaml, recipe, B
        P = self.P = iw.Problem(yaml, overrides=ov)
        bt = self.bt = iw.Batch(P, local, B)
        self.n, self.me, self.mi, self.rows = P.qp_shape()
        self.masks, self.beh = None, None
        self.buckets = [(0, B, P.default_contact_mask)]
        if recipe == "walk":
            P.set_task_weight("momentum", 0.0)  # BEHAVIOR.customize_task_weights, etc/talos/walk.yaml:8-9
            self.q, self.v, self.refs, self.masks, self.phase, self.beh = make_walk_workload(P, bt, B, 0xB200 + 5 + 1000 * rank, local)
        else:
            self.q, self.v, self.refs = make_inputs(P, B, recipe, 0xB200 + 3 + 1000 * rank)
        if recipe == "walk_standin":
            P.set_task_weight("momentum", 0.0)
            self.masks = make_walk_masks(P, B, self.refs, 0xB200 + 5 + 1000 * rank)
            self.buckets, first = [], 0
            for m in (3, 2, 1):
                cnt = int((self.masks == m).sum())
                if cnt:
                    self.buckets.append((first, cnt, m))
                first += cnt
        self.dq, self.dv, self.dr = (torch.from_numpy(x).to(dev) for x in (self.q, self.v, self.refs))
        f64 = dict(dtype=torch.float64, device=dev)
        self.out = {"tau": torch.empty(B, P.na, **f64), "dv": torch.empty(B, P.nv, **f64), "f": torch.empty(B, max(12 * P.ncontacts, 1), **f64),
                    "q_next": torch.empty(B, P.nq, **f64), "v_next": torch.empty(B, P.nv, **f64),
                    "status": torch.empty(B, dtype=torch.int32, device=dev), "iters": torch.empty(B, dtype=torch.int32, device=dev)}
        self.ptrs = {k: t.data_ptr() for k, t in self.out.items()}
        self.esz = {"tau": 8 * P.na, "dv": 8 * P.nv, "f": 8 * max(12 * P.ncontacts, 1), "q_next": 8 * P.nq, "v_next": 8 * P.nv, "status": 4, "iters": 4}
        self.stream = torch.cuda.Stream(device=dev)  # an explicit stream: the timing events must see the kernels
        if self.beh is not None:
            self.d_phase = torch.from_numpy(self.phase).to(dev)
            self.d_maskbuf = torch.zeros(B, dtype=torch.int32, device=dev)
            self.walk_tick = 0
        # end-to-end arm: pinned host buffers and results
        self.hq, self.hv, self.hr = (torch.from_numpy(x).pin_memory() for x in (self.q, self.v, self.refs))
        self.res = iw.Results(P, B, fields=("tau", "dv", "f", "q_next", "v_next", "status"))
        self.pinned = {}
        for fld in ("tau", "dv", "f", "q_next", "v_next", "status"):
            t = torch.from_numpy(getattr(self.res, fld)).pin_memory()
            self.pinned[fld] = t
            setattr(self.res, fld, t.numpy())

    def step(self):
        P, bt, B = self.P, self.bt, self.B
        if self.beh is not None:
            # closed device-side tick: behaviour -> refs + contact masks, device-side bucketing, one launch set per mask, state
            # integrated in place (every timed step is the NEXT tick of every instance's walk)
            bt.behavior_rollout_device(self.beh, B, self.dq.data_ptr(), self.dv.data_ptr(), self.dr.data_ptr(), self.d_phase.data_ptr(), self.walk_tick, 1,
                                       init_refs=False, out_ptrs=self.ptrs, mask_ptr=self.d_maskbuf.data_ptr(), stream=self.stream.cuda_stream)
            self.walk_tick += 1
            return
        for first, cnt, m in self.buckets:  # one shape-uniform launch set per contact-mask bucket (instances are ordered by mask)
            bt.update_device(cnt, self.dq.data_ptr() + 8 * P.nq * first, self.dv.data_ptr() + 8 * P.nv * first, self.dr.data_ptr() + 8 * P.nref * first,
                             P.nref, m, {k: p + self.esz[k] * first for k, p in self.ptrs.items()}, self.stream.cuda_stream)

    def host_call(self):
        self.bt.update(self.hq.numpy(), self.hv.numpy(), self.hr.numpy(), contact_mask=self.masks, out=self.res)  # synchronous

    def io_bytes(self):
        h2d = self.q.nbytes + self.v.nbytes + self.refs.nbytes
        d2h = sum(getattr(self.res, f).nbytes for f in ("tau", "dv", "f", "q_next", "v_next", "status"))
        return h2d, d2h


def run_workload(ctx, workload, steps, warmup, batch=0, with_cpu=False, peak=None):
    """One bench line (dict, on rank 0; None elsewhere) for `workload` on every rank's GPU."""
    import torch
    import inria_wbc_b200 as iw
    rank, world, local, dev, dist = ctx["rank"], ctx["world"], ctx["local"], ctx["dev"], ctx["dist"]
    if workload in MIXED:
        a, b, each = MIXED[workload]
        parts = [Part(iw, torch, a, batch // 2 if batch else each, rank, local, dev), Part(iw, torch, b, batch // 2 if batch else each, rank, local, dev)]
    else:
        parts = [Part(iw, torch, workload, batch or WORKLOADS[workload][2], rank, local, dev)]
    B = sum(p.B for p in parts)
    flush = ctx["flush"]
    main = parts[0].stream
    fork, joins = torch.cuda.Event(), [torch.cuda.Event() for _ in parts[1:]]

    def step():
        # the parts (robots) of a mixed batch run concurrently on their own streams, bracketed by the first part's stream
        if len(parts) > 1:
            fork.record(main)
        parts[0].step()
        for p, j in zip(parts[1:], joins):
            p.stream.wait_event(fork)
            p.step()
            j.record(p.stream)
            main.wait_event(j)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    launches0 = sum(p.bt.launch_count for p in parts)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for k in range(steps):
        with torch.cuda.stream(main):
            flush.fill_(float(k))  # L2 flush, outside the timed events
            ev[k][0].record(main)
            step()
            ev[k][1].record(main)
    barrier()
    launches = sum(p.bt.launch_count for p in parts) - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    clocks = sampler.stop() if rank == 0 else None
    n_opt = sum(int((p.out["status"].cpu().numpy() == 0).sum()) for p in parts)
    it_sum = [float(p.out["iters"].cpu().numpy().sum()) for p in parts]

    # ---- end-to-end arm: pinned host buffers through the reference-facing call (H2D + kernels + D2H inside the timed region)
    e2e_steps = max(3, steps)
    for _ in range(3):
        for p in parts:
            p.host_call()
    barrier()
    e2e_lat = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        for p in parts:
            p.host_call()
        e2e_lat.append(time.perf_counter() - t1)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = sum(p.io_bytes()[0] for p in parts)
    d2h = sum(p.io_bytes()[1] for p in parts)

    # ---- max over ranks
    tt = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    agg = torch.tensor([float(n_opt)] + it_sum, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_ms, e2e_s = tt.tolist()
    agg = agg.tolist()
    n_opt_all, it_all = agg[0], agg[1:]
    if rank != 0:
        return None
    ms_per_step = total_ms / steps
    value = B * world / (ms_per_step * 1e-3)
    e2e_val = B * world / (e2e_s / e2e_steps)
    kern_ms = float(np.mean(step_ms))
    flops_step, fl_parts, io_step = 0.0, {}, 0
    for p, its in zip(parts, it_all):
        P = p.P
        adds = max(its / (p.B * world) - 1.0, 0.0)
        nj, nfr, nsc = problem_counts(P)
        fl = algorithmic_flops(P.nv, P.na, p.n, p.me, p.mi, p.rows, adds, nj, nfr, nsc)
        fl["source"] = "F_qp: SURVEY.md 8(d) reference-algorithm formula with the measured mean additions; F_dyn, F_asm: textbook estimate"
        try:  # F_dyn / F_asm counted by the instrumented oracle (tools/count_flops.py, oracle/count_scalar.hpp)
            counted = json.load(open(os.path.join(ROOT, "profiles", "flops_per_update.json"))).get(p.workload)
            if counted:
                fl["f_dyn"], fl["f_asm"] = counted["f_dyn"], counted["f_asm"]
                fl["total"] = fl["f_qp"] + fl["f_dyn"] + fl["f_asm"]
                fl["f_qp_executed_by_oracle"] = counted["f_qp_executed_by_oracle"]
                fl["source"] = ("F_qp: SURVEY.md 8(d) reference-algorithm formula with the measured mean additions; F_dyn, F_asm: counted by the "
                                "instrumented oracle (profiles/flops_per_update.json)")
        except Exception:
            pass
        fl["mean_active_set_iterations"] = its / (p.B * world)
        fl_parts[p.workload] = fl
        flops_step += fl["total"] * p.B
        io_step += (8 * (P.nq + P.nv + P.nref) + 8 * (P.na + 2 * P.nv + 12 * P.ncontacts + P.nq) + 8) * p.B
    achieved = flops_step / (kern_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload, {}).get("bytes_per_launch")
    except Exception:
        pass
    p0 = parts[0]
    line = {
        "metric": "batched TSID control updates (QP solves) per second", "value": value, "unit": "updates/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "controller": [p.yaml for p in parts] if len(parts) > 1 else p0.yaml, "per_gpu_batch": B,
                   "qp": {"n": p0.n, "me": p0.me, "mi": p0.mi, "cost_rows": p0.rows} if len(parts) == 1 else
                   {p.workload: {"n": p.n, "me": p.me, "mi": p.mi, "cost_rows": p.rows, "batch": p.B} for p in parts},
                   "l2": "flushed (256 MB write) between timed steps", "optimal_fraction": n_opt_all / (B * world),
                   "mean_active_set_iterations": sum(it_all) / (B * world), "p50_device_ms_per_batch": float(np.median(step_ms)),
                   "p95_device_ms_per_batch": float(np.percentile(step_ms, 95)),
                   "parallelism": "dp%d (independent shards, no collective)" % world,
                   "states": ("random ticks of humanoid::walk (etc/talos/walk.yaml): a seed instance is walked on the device, instances are its "
                              "snapshots + noise; every timed step is the next tick, references and contact masks generated on the device")
                   if p0.beh is not None else "SURVEY.md 8(d) randomized-state recipe " + ", ".join("'%s'" % p.recipe for p in parts)},
        "e2e": {"value": e2e_val, "unit": "updates/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s / e2e_steps, "steps": e2e_steps,
                "p50_cycle_latency_ms": 1e3 * float(np.percentile(e2e_lat, 50)), "p95_cycle_latency_ms": 1e3 * float(np.percentile(e2e_lat, 95))},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "bound_note": "FP64 FMA on CUDA cores + shared-memory bandwidth; neither hbm nor tensor bound (SURVEY.md 8(d))",
                     "achieved": achieved, "peak": peak["measured_tflops"], "unit": "TFLOP/s",
                     "frac": achieved / peak["measured_tflops"] if peak["measured_tflops"] > 0 else None,
                     "peak_source": "measured live (dependent-FMA probe, iwbc_measure_fp64_tflops); MEASURED_PEAKS.json holds no FP64 figure",
                     "peak_probe": peak, "frac_of_clock_bound": achieved / peak["clock_bound_tflops"] if peak.get("clock_bound_tflops") else None,
                     "flops_per_update": fl_parts[p0.workload] if len(parts) == 1 else fl_parts, "kernel_ms": kern_ms, "traffic": traffic,
                     # measured DRAM bytes per launch (ncu, profiles/traffic.json) over this run's kernel time: what the kernel really asks of
                     # HBM -- for the thread-per-instance arm kernel (local-memory working set) this, not FP64, is the nearer roof
                     "traffic_hbm": ({"achieved": traffic * (B / WORKLOADS[workload][2] if workload in WORKLOADS else 1.0) / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak,
                                      "unit": "GB/s", "frac": traffic * (B / WORKLOADS[workload][2] if workload in WORKLOADS else 1.0) / (kern_ms * 1e-3) / 1e9 / hbm_peak}
                                     if traffic else None),
                     "hbm": {"achieved": io_step / (kern_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": io_step / (kern_ms * 1e-3) / 1e9 / hbm_peak, "bytes_per_update": io_step / B,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}},
        "clocks": clocks,
    }
    if with_cpu and len(parts) == 1:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import problem_loader
        Pc = problem_loader.load(p0.yaml, overrides=WORKLOADS[workload][1])  # the oracle's own reader of the same configuration
        threads = os.cpu_count() or 1
        ns = min(p0.B, 4096)
        rate, done, el = oracle_rate(Pc, p0.q[:ns], p0.v[:ns], p0.refs[:ns], 10.0, threads)
        rate1, done1, el1 = oracle_rate(Pc, p0.q[:256], p0.v[:256], p0.refs[:256], 2.0, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "updates/s", "cores": threads, "kind": "port",
                                "sample": "%d updates of the same workload in %.2f s on %d threads (oracle port)" % (done, el, threads),
                                "single_thread": rate1,
                                "reference_recorded": "0.0126 ms/cycle Franka, different unknown hardware (tests/ref_test_franka.yaml:11)"}
    return line


def fp64_peak(iw, local):
    """The roofline denominator: a live dependent-FMA probe, with the clock record it was taken at and the bound from the SM count."""
    sampler = ClockSampler(local)
    sampler.start()
    vals = [iw.measure_fp64_tflops(local) for _ in range(3)]
    clk = sampler.stop()
    import torch
    sms = torch.cuda.get_device_properties(local).multi_processor_count
    mhz = clk.get("sm_max_mhz") or 1965.0
    return {"measured_tflops": max(vals), "runs": vals, "clocks": clk, "sm_count": sms, "fp64_fma_lanes_per_sm": 64,
            "clock_bound_tflops": sms * 64 * 2 * mhz * 1e6 / 1e12,
            "how": "8 independent dependent-FMA chains per thread, 8 CTAs x 256 threads per SM, best of 4 launches, CUDA events"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS) + sorted(MIXED))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs appended to the default line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import inria_wbc_b200 as iw
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the batched cycle has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    try:  # N > 1: run this rank (and allocate its pinned buffers) on the CPUs next to its GPU: the host-buffer arm moves ~24 GB/s per GPU
        if world <= 1:
            raise RuntimeError("single rank: keep every core for the CPU-baseline leg")
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
    except Exception:
        pass
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    ctx = {"rank": rank, "world": world, "local": local, "dev": dev, "dist": dist,
           "flush": torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)}  # > 126 MB L2
    peak = fp64_peak(iw, local)
    line = run_workload(ctx, args.workload, args.steps, args.warmup, batch=args.batch, with_cpu=(world == 1 and not args.no_cpu), peak=peak)
    if args.workload == DEFAULT_WORKLOAD and not args.no_secondary and not args.batch:
        # BASELINE.json's other configurations, measured by the same program right after on the same GPUs (every rank its own shard, at
        # every N: configs[4] = Talos walking is 131 072 states per GPU, i.e. 1 M over 8): configs[1] Franka batch, configs[3] iCub /
        # Tiago alone and as ONE mixed step, configs[4] walking
        sec = {}
        for wl in SECONDARY:
            try:
                d = run_workload(ctx, wl, 10, 3, peak=peak)
                if d is not None:
                    sec[wl] = {"value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "e2e": d["e2e"]["value"],
                               "p50_cycle_latency_ms": d["e2e"].get("p50_cycle_latency_ms"), "roofline_frac": d["roofline"]["frac"],
                               "traffic_hbm_frac": (d["roofline"].get("traffic_hbm") or {}).get("frac"),
                               "gpu_launches": d["gpu_launches"], "per_gpu_batch": d["config"]["per_gpu_batch"], "n_gpus": d["n_gpus"],
                               "optimal_fraction": d["config"]["optimal_fraction"]}
            except Exception as e:  # the headline line must not depend on it
                if rank == 0:
                    sec[wl] = {"error": str(e)[:200]}
            torch.cuda.empty_cache()
        if line is not None:
            line["secondary"] = sec
    if line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
