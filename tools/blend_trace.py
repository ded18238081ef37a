"""Per-warp timeline of the forward blend (diagnostic; needs the -DSGS_BLEND_TRACE build:
tools/variants.sh + SGS_LIB_PATH).  Prints how the kernel's time is spent: per-SM busy time, the
warps that finish last, the relation between a warp's list length / relevant pairs and its time."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from sings_b200 import _lib
from sings_b200.step import AvatarStep, FrameInputs

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda", 0)
av, pose, transl, view, G, bg = bench.build_frame_inputs(cfg, 0)
t = lambda a: torch.as_tensor(a, device=dev)
st = AvatarStep(t(av.xyz_canon), None if cfg["iso"] else t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs),
                t(av.lbs_weights), t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano), cfg["H"], cfg["W"], cfg["D"])
fr = FrameInputs(pose=t(pose), transl=t(transl), viewmatrix=t(view.world_view_transform), projmatrix=t(view.full_proj_transform),
                 campos=t(view.camera_center), bg=t(bg), tanfovx=view.tanfovx, tanfovy=view.tanfovy)
for _ in range(5):
    st.forward(fr); st.backward(t(G))
torch.cuda.synchronize()
L = _lib.lib()
L.sgs_debug_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
tiles = ((cfg["H"] + 15) // 16) * ((cfg["W"] + 15) // 16)
n = min(tiles * 8, 1 << 16)
dt = np.dtype([("t0", "<u8"), ("t1", "<u8"), ("smid", "<u4"), ("tile", "<u4"), ("warp", "<u4"), ("rel", "<u4"), ("len", "<u4"), ("head", "<u4")])
buf = np.zeros(n, dt)
rc = L.sgs_debug_trace_read(buf.ctypes.data, buf.nbytes)
assert rc == 0, rc
b = buf[buf["t1"] > 0]
T0 = b["t0"].min()
t0 = (b["t0"] - T0) / 1e3; t1 = (b["t1"] - T0) / 1e3
dur = t1 - t0
end = t1.max()
print(f"warps {len(b)}, kernel span {end:.1f} us, sum of warp durations {dur.sum():.0f} us, non-empty warps {(b['rel'] > 0).sum()}")
print(f"relevant pairs total {b['rel'].sum()}, list entries scanned total {b['len'].sum()}")
# per SM: first start, last end, busy warps-time
sm = b["smid"]
ends = np.array([t1[sm == s].max() for s in np.unique(sm)])
print("per-SM last end (us): min %.1f  p25 %.1f  median %.1f  p75 %.1f  max %.1f" % (ends.min(), *np.percentile(ends, [25, 50, 75]), ends.max()))
# resident-warp count over time
ev = np.concatenate([np.stack([t0, np.ones_like(t0)], 1), np.stack([t1, -np.ones_like(t1)], 1)])
ev = ev[np.argsort(ev[:, 0])]
cum = np.cumsum(ev[:, 1])
for frac in (0.1, 0.25, 0.5, 0.6, 0.7, 0.8, 0.9, 0.95):
    tt = frac * end
    k = np.searchsorted(ev[:, 0], tt) - 1
    print(f"  t = {tt:6.1f} us ({frac:.2f}): {int(cum[k])} warps resident ({cum[k] / 148:.1f} per SM)")
# the last finishers
order = np.argsort(-t1)[:12]
print("last finishers: end us, start us, dur us, tile, warp, list len, relevant, us per relevant pair")
for i in order:
    print(f"  {t1[i]:7.1f} {t0[i]:7.1f} {dur[i]:7.1f}  tile {b['tile'][i]:5d} w{b['warp'][i]}  len {b['len'][i]:5d} rel {b['rel'][i]:5d}  {dur[i] / max(1, b['rel'][i]) * 1e3:.1f} ns")
# time per relevant pair by start-time decile
nz = b["rel"] > 16
q = np.percentile(t0[nz], np.arange(0, 101, 10))
print("start-time decile: warps, mean duration, ns per relevant pair, ns per list entry")
for lo, hi in zip(q[:-1], q[1:]):
    m = nz & (t0 >= lo) & (t0 <= hi)
    print(f"  [{lo:6.1f},{hi:6.1f}] {m.sum():5d}  {dur[m].mean():6.1f} us  {dur[m].sum() / b['rel'][m].sum() * 1e3:6.1f}  {dur[m].sum() / b['len'][m].sum() * 1e3:6.1f}")
long_ = np.argsort(-b["rel"])[:8]
print("longest chains: rel, len, dur, start")
for i in long_:
    print(f"  rel {b['rel'][i]:5d} len {b['len'][i]:5d} dur {dur[i]:6.1f} start {t0[i]:6.1f}")
