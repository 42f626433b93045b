"""Batched reference generators: the callers of the hot path (host side for now).

Re-hosts the two behaviors the reference's parity tests use, as *tables of reference vectors*:
  generic::cartesian   src/behaviors/generic/cartesian.cpp:8-94   (SE3 min-jerk, optional loop)
  humanoid::move_com   src/behaviors/humanoid/move_com.cpp:8-61   (CoM min-jerk, optional loop)
`table(ticks)` returns refs[ticks][nref]: row t is what the reference's behavior writes into the
tasks before the controller update of tick t.
"""
import numpy as np
import yaml

from . import trajs


def _load(path_or_dict):
    if isinstance(path_or_dict, dict):
        return path_or_dict
    return yaml.safe_load(open(path_or_dict))


class Cartesian:
    def __init__(self, problem, config, feed_forward=True):
        c = _load(config)["BEHAVIOR"]
        self.P = problem
        self.dt = problem.struct().dt
        dur = float(np.float32(c["trajectory_duration"]))
        self.loop = bool(c["loop"])
        names = c["task_names"]
        ts, to = c["relative_targets_pos"], c["relative_targets_rpy"]
        if len(names) != len(ts):
            raise ValueError("cartesian behavior needs the same number of tasks and targets")
        ref0 = problem.default_refs()
        self.ref0 = ref0
        self.segments = []  # per task: list of (pos, rot, vel, acc) tables
        self.slots = []
        for i, name in enumerate(names):
            off, _ = problem.ref_slot(name)
            p0, R0 = trajs.sample_se3(ref0[off:off + 24])
            pf, Rf = p0.copy(), R0.copy()
            if len(ts[i]) == 3:
                pf = p0 + np.asarray(ts[i], dtype=np.float64)
            if i < len(to) and len(to[i]) == 3:
                Rf = trajs.rpy_matrix(to[i]) @ R0
            segs = [trajs.min_jerk_se3(p0, R0, pf, Rf, self.dt, dur)]
            if self.loop:
                segs.append(trajs.min_jerk_se3(pf, Rf, p0, R0, self.dt, dur))
            self.segments.append(segs)
            self.slots.append(off)
        self.ff = feed_forward

    def table(self, ticks):
        refs = np.tile(self.ref0, (ticks, 1))
        nseg = len(self.segments[0])
        seglen = len(self.segments[0][0][0])
        for t in range(ticks):
            sel = t // seglen
            if self.loop:
                sel %= nseg
            if sel >= nseg:
                # finished and not looping: the last written reference stays
                for off in self.slots:
                    refs[t, off:off + 24] = refs[t - 1, off:off + 24]
                continue
            i = t % seglen
            for k, off in enumerate(self.slots):
                pos, rot, vel, acc = self.segments[k][sel]
                refs[t, off:off + 24] = trajs.se3_sample(pos[i], rot[i], vel[i] if self.ff else None, acc[i] if self.ff else None)
        return refs


class MoveCom:
    def __init__(self, problem, config):
        c = _load(config)["BEHAVIOR"]
        self.P = problem
        self.dt = problem.struct().dt
        dur = float(np.float32(c["trajectory_duration"]))
        self.loop = bool(c["loop"])
        absolute = bool(c["absolute"])
        mask = str(c["mask"])
        if len(mask) != 3:
            raise ValueError("The mask for the CoM should be 3-dimensional")
        ref0 = problem.default_refs()
        self.ref0 = ref0
        self.off, _ = problem.ref_slot("com")
        init = ref0[self.off:self.off + 3].copy()
        targets = [list(t) for t in c["targets"]]
        if self.loop:
            targets.append(list(init) if absolute else [0.0, 0.0, 0.0])
        self.segs = []
        start = init.copy()
        for tgt in targets:
            if len(tgt) != 3:
                raise ValueError("references need to be 3-dimensional")
            end = init.copy()
            for k in range(3):
                if mask[k] == "1":
                    end[k] = tgt[k] if absolute else tgt[k] + init[k]
            self.segs.append(tuple(trajs.min_jerk_vec(start, end, self.dt, dur, o) for o in (0, 1, 2)))
            start = end

    def table(self, ticks):
        refs = np.tile(self.ref0, (ticks, 1))
        pos = np.concatenate([s[0] for s in self.segs])
        vel = np.concatenate([s[1] for s in self.segs])
        acc = np.concatenate([s[2] for s in self.segs])
        n = len(pos)
        for t in range(ticks):
            i = t % n if self.loop else min(t, n - 1)
            refs[t, self.off:self.off + 3] = pos[i]
            refs[t, self.off + 3:self.off + 6] = vel[i]
            refs[t, self.off + 6:self.off + 9] = acc[i]
        return refs
