"""Cost of the per-stage event-record nodes inside the frame graph: frames/s of the same
CUDA-graph frame captured with and without stage timing (run under gpurun)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from sings_b200.step import AvatarStep, FrameInputs

dev = torch.device("cuda", 0)
t = lambda a: torch.as_tensor(a, device=dev)
res = {}
for timing in (True, False, True, False):
    sets = []
    for r in range(bench.RING):
        av, pose, transl, view, G, bg = bench.build_frame_inputs(10 * r)
        st = AvatarStep(t(av.xyz_canon), t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs), t(av.lbs_weights),
                        t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano), bench.H_IMG, bench.W_IMG,
                        bench.SH_DEG, timing=timing)
        fr = FrameInputs(pose=t(pose), transl=t(transl), viewmatrix=t(view.world_view_transform),
                         projmatrix=t(view.full_proj_transform), campos=t(view.camera_center), bg=t(bg),
                         tanfovx=view.tanfovx, tanfovy=view.tanfovy)
        Gd = t(G)
        for _ in range(3):
            st.forward(fr); st.backward(Gd)
        torch.cuda.synchronize()
        st.check_capacity()
        sets.append((st, st.capture(fr, Gd), fr, Gd))       # the graph binds fr and Gd by address: keep them alive
    for i in range(20):
        sets[i % bench.RING][1]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 400
    e0.record()
    for i in range(n):
        sets[i % bench.RING][1]()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"timing={timing}: {ms*1000:.1f} us/frame  {1000/ms:.1f} frames/s")
    del sets
