"""CPU restatement of the reference's scale-edge loss and of the K-NN it calls (TEST INFRASTRUCTURE:
only tests/ and bench.py's comparison legs may import it).

  knn_points      pytorch3d.ops.knn_points -- THIRD-PARTY, not vendored under /root/reference and not
                  installed here (requirements: pytorch3d, unpinned): its published behaviour is
                  restated -- exact K nearest neighbours by squared Euclidean distance, sorted
                  ascending, returned as (dists, idx) -- by brute force.  PARITY UNPINNED for this
                  function (no golden vectors can be produced without the package); what anchors
                  it is its definition: tests compare against this brute force.
  GaussiansEdgeLoss.forward    /root/reference/sings/rec/losses/loss_items.py:57-90, statement by statement.
                  PINNED around knn_points: tests/golden/reg_golden_edge.npz holds the loss and its gradient as the
                  reference's own class computes them when knn_points is this brute force
                  (tests/golden/make_reg_golden.py); tests/test_reg_oracle.py checks gaussians_edge_loss against it.
"""
import torch


def knn_points(p1: torch.Tensor, p2: torch.Tensor, K: int):
    """(1, N, 3), (1, M, 3) -> dists (1, N, K) squared, idx (1, N, K), ascending (pytorch3d semantics)."""
    d = torch.cdist(p1[0].double(), p2[0].double()) ** 2
    dists, idx = torch.topk(d, K, dim=1, largest=False, sorted=True)
    return dists.to(p1.dtype)[None], idx[None], None


def gaussians_edge_loss(human_gs_out, K: int = 9):
    verts = human_gs_out["xyz_canon"]
    scales = human_gs_out["scales"][:, 0]
    dists_knn, idx_knn, _ = knn_points(verts.unsqueeze(0), verts.unsqueeze(0), K=K)
    edge_vectors = verts[idx_knn[0, :, 1:]] - verts.unsqueeze(1)
    edge_lengths = torch.norm(edge_vectors, dim=-1).mean(dim=-1, keepdim=True).detach()
    scale_proj_i = scales.unsqueeze(1)
    len_factor = 1.0
    return ((scale_proj_i - len_factor * edge_lengths) ** 2).mean()
