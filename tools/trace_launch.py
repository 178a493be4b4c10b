"""Reads the -DZS_TRACE stamps of one K-step launch (tools/trace_launch.sh): per-warp global-timer values at kernel
entry, after staging / load_state / build_grid, after each of the first steps and at exit."""
import ctypes as C
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np
import torch
import parity_util as pu
from libzombsole_b200 import abi, _native
from libzombsole_b200.engine import ZsEngine

W, S = 8192, 40


def grab(L, clear=True):
    buf = np.zeros((W, S), np.uint64)
    L.zs_debug_trace.argtypes = [C.c_void_p, C.c_int32]
    rc = L.zs_debug_trace(buf.ctypes.data, 1 if clear else 0)
    assert rc == 0
    return buf


def q(a):
    return "min %7.2f  p50 %7.2f  p90 %7.2f  p99 %7.2f  max %7.2f" % tuple(np.percentile(a, [0, 50, 90, 99, 100]) / 1e3)


def main():
    Ks = [int(a) for a in sys.argv[1:]] or [1, 4, 20]
    N = 4096
    cfg, m = pu.build(pu.CONFIGS["c1_bridge_ext"], N, 0, auto_reset=True, max_episode_steps=1000)
    eng = ZsEngine(cfg, m)
    L = _native.lib()
    obs = eng.new_obs(13)
    KM = max(Ks + [200])
    rew, term, trunc = eng.new_outputs(KM)
    acts = torch.zeros((KM, N, 1), dtype=torch.int32, device=eng.device)
    for s in range(KM):
        eng.fill_synthetic_actions(s, acts[s])
    eng.rollout(200, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
    torch.cuda.synchronize()
    for K in Ks:
        for _ in range(5):  # back to back; the trace keeps the last launch
            eng.rollout(K, 0, acts, abi.ACTIONS_DISCRETE, obs, rew, term, trunc)
        torch.cuda.synchronize()
        t = grab(L)
        used = t[:, 0] > 0
        t = t[used].astype(np.int64)
        t0 = t[:, 0].min()
        print("K=%d: %d warps on %d SMs; all times in us" % (K, len(t), len(set(t[:, 31].tolist()))))
        print("  entry after first entry   ", q(t[:, 0] - t0))
        print("  staging                   ", q(t[:, 1] - t[:, 0]))
        print("  load_state                ", q(t[:, 2] - t[:, 1]))
        print("  build_grid                ", q(t[:, 3] - t[:, 2]))
        print("  wait for the obs template  ", q(t[:, 29] - t[:, 3]))
        prev = t[:, 29]
        for s in range(min(K, 16)):
            cur = t[:, 4 + s]
            print("  step %2d                   " % s, q(cur - prev))
            prev = cur
        print("  last traced step -> exit  ", q(t[:, 30] - prev))
        last0 = t[:, 4 + K - 2] if 2 <= K <= 16 else t[:, 29]
        if K <= 16:
            print("  LAST step: template issue  ", q(t[:, 24] - last0))
            print("  LAST step: world_step      ", q(t[:, 25] - t[:, 24]))
            print("  LAST step: reward/rules/out", q(t[:, 26] - t[:, 25]))
            print("  LAST step: world init      ", q(t[:, 27] - t[:, 26]), " (%d warps > 1 us)" % int(((t[:, 27] - t[:, 26]) > 1000).sum()))
            print("  LAST step: obs patches     ", q(t[:, 4 + K - 1] - t[:, 27]))
        if K <= 16:  # what a world init costs the step it happens in, phase by phase
            ini = (t[:, 27] - t[:, 26]) > 1000
            if ini.any() and (~ini).any():
                ph = [("template issue", t[:, 24] - last0), ("world_step", t[:, 25] - t[:, 24]), ("reward/rules/out", t[:, 26] - t[:, 25]),
                      ("world init", t[:, 27] - t[:, 26]), ("obs patches", t[:, 4 + K - 1] - t[:, 27]), ("whole step", t[:, 4 + K - 1] - last0)]
                print("  LAST step, warps with a world init (%d) against the others, mean us: " % int(ini.sum()) +
                      ", ".join("%s %.2f / %.2f" % (nm, v[ini].mean() / 1e3, v[~ini].mean() / 1e3) for nm, v in ph))
        did = t[:, 23] > t[:, 20]
        did &= t[:, 20] > t0
        if did.any():
            print("  world init (last one of %d warps): draws+shuffle %s" % (int(did.sum()), q((t[:, 21] - t[:, 20])[did])))
            print("                                     grid/lists    %s" % q((t[:, 22] - t[:, 21])[did]))
            print("                                     placement     %s" % q((t[:, 23] - t[:, 22])[did]))
        print("  exit after first entry    ", q(t[:, 30] - t0), "  (kernel span %.2f us)" % ((t[:, 30].max() - t0) / 1e3))
        if K <= 16:  # is the slow tail of the first step made of each SM's FIRST warps (cold instruction / data caches)?
            ws = t[:, 25] - t[:, 24]
            smv = t[:, 31]
            by_rank = {}
            for i in set(smv.tolist()):
                idx = np.nonzero(smv == i)[0]
                order = idx[np.argsort(t[idx, 0], kind="stable")]
                for r, w in enumerate(order):
                    by_rank.setdefault(r, []).append(ws[w])
            print("  LAST step world_step by entry order within the SM (mean us):",
                  {r: round(float(np.mean(v)) / 1e3, 2) for r, v in sorted(by_rank.items())})
            slow = ws > 2 * np.median(ws)
            print("  slow warps (> 2 x median): %d of %d; on %d SMs" % (int(slow.sum()), len(ws), len(set(smv[slow].tolist()))))
            names = ["wander", "idle", "hits", "static hit", "sequential execute", "deaths", "re-rank", "fresh world", "world init"]
            fl = t[:, 28]
            for b, nm in enumerate(names):
                took = (fl >> b) & 1 == 1
                if took.any():
                    print("    path %-20s taken by %4d warps: world_step mean %.2f us (others %.2f); %d of the slow warps" % (
                        nm, int(took.sum()), ws[took].mean() / 1e3, ws[~took].mean() / 1e3 if (~took).any() else 0, int((took & slow).sum())))
        # every warp's slowest step of the launch, and which of the less common paths it took in that step
        names = ["wander", "idle", "hits", "static hit", "sequential execute", "deaths", "re-rank", "fresh world", "world init"]
        dmax, fmax = t[:, 32], t[:, 33]
        print("  slowest step of a warp       ", q(dmax))
        slowest = dmax > np.percentile(dmax, 90)
        for b, nm in enumerate(names):
            took = (fmax >> b) & 1 == 1
            if took.any():
                print("    its path %-20s in %4d warps: that step mean %6.2f us (others %6.2f); %3d of the %d warps with the slowest tenth" % (
                    nm, int(took.sum()), dmax[took].mean() / 1e3, dmax[~took].mean() / 1e3 if (~took).any() else 0,
                    int((took & slowest).sum()), int(slowest.sum())))
        combo = {}
        for f, d in zip(fmax[slowest].tolist(), dmax[slowest].tolist()):
            key = "+".join(nm for b, nm in enumerate(names) if (f >> b) & 1 and nm not in ("hits", "re-rank", "deaths")) or "-"
            combo.setdefault(key, []).append(d)
        for key, v in sorted(combo.items(), key=lambda kv: -len(kv[1]))[:8]:
            print("    slowest tenth: %4d warps took {%s}: mean %.2f us" % (len(v), key, np.mean(v) / 1e3))
        # per SM: warps, span
        sm = t[:, 31]
        per = [(int(i), int((sm == i).sum()), (t[sm == i, 30].max() - t0) / 1e3) for i in sorted(set(sm.tolist()))]
        cnt = np.array([p[1] for p in per]); end = np.array([p[2] for p in per])
        print("  warps per SM: min %d max %d; SM finish time: min %.2f p50 %.2f max %.2f" % (cnt.min(), cnt.max(), end.min(), np.median(end), end.max()))
        print("  finish time by warps-per-SM:", {int(c): round(float(end[cnt == c].mean()), 2) for c in sorted(set(cnt.tolist()))})
    eng.close()


if __name__ == "__main__":
    main()
