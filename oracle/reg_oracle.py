"""CPU restatement of the reference's non-image loss terms (TEST INFRASTRUCTURE: only tests/ and
bench.py's comparison legs may import it; the product path never does).

Follows, statement by statement,
  /root/reference/sings/rec/losses/loss_items.py:15-54     L2Norm
  /root/reference/sings/rec/losses/loss_items.py:93-190    RegionLaplacianLoss_v2 ("standard" branch:
                                                           reset_laplacians :126-156, forward_hands :163-171, forward :174-190)
  /root/reference/sings/rec/losses/loss_items.py:194-214   build_edges, pcd_laplacian_smoothing
  /root/reference/sings/rec/utils/body_model/smpl_parsing.py:38-44   parse_weights
  laplacian       pytorch3d.ops.laplacian -- THIRD-PARTY, not vendored under /root/reference and not
                  installed here (install_all.sh:21 installs pytorch3d's default branch, unpinned).  Its published
                  algorithm (pytorch3d/ops/laplacian_matrices.py) is restated with dense matrices:
                  A[e0, e1] += 1, A[e1, e0] += 1; deg = row sums; L = A / deg (rows with deg 0 stay 0); L -= I
                  (`laplacian_sparse`: the same with torch.sparse COO tensors, its own construction, for large sizes).
PINNED for L2Norm and for everything RegionLaplacianLoss_v2 / pcd_laplacian_smoothing do AROUND that
function: tests/golden/reg_golden_*.npz hold outputs and gradients of the reference's own classes run
with this `laplacian` standing in for pytorch3d's (tests/golden/make_reg_golden.py);
tests/test_reg_oracle.py checks this file against them.  PARITY UNPINNED for `laplacian` itself.
"""
import numpy as np
import torch


def laplacian(verts: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """Dense (V, V) float32 matrix with the values of pytorch3d.ops.laplacian(verts, edges)."""
    V = verts.shape[0]
    e0, e1 = edges.long().unbind(1)
    A = torch.zeros(V, V, dtype=torch.float32)
    A.index_put_((e0, e1), torch.ones(e0.numel()), accumulate=True)
    A.index_put_((e1, e0), torch.ones(e0.numel()), accumulate=True)
    deg = A.sum(dim=1)
    inv = torch.where(deg > 0.0, 1.0 / deg, deg)
    L = torch.zeros(V, V, dtype=torch.float32)
    # one entry of value inv[row] per (directed) occurrence of the edge, summed like a COO tensor
    L.index_put_((e0, e1), inv[e0], accumulate=True)
    L.index_put_((e1, e0), inv[e1], accumulate=True)
    L -= torch.eye(V)
    return L


def laplacian_sparse(verts: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """The same operator as a torch sparse COO tensor on verts' device, built the way pytorch3d builds it (for
    sizes where the dense matrix does not fit: the large GPU test and bench.py's eager-torch comparator)."""
    V, dev = verts.shape[0], verts.device
    e0, e1 = edges.long().unbind(1)
    idx = torch.cat([torch.stack([e0, e1], dim=1), torch.stack([e1, e0], dim=1)], dim=0).t()
    ones = torch.ones(idx.shape[1], dtype=torch.float32, device=dev)
    A = torch.sparse_coo_tensor(idx, ones, (V, V))
    deg = torch.sparse.sum(A, dim=1).to_dense()
    deg0, deg1 = deg[e0], deg[e1]
    deg0 = torch.where(deg0 > 0.0, 1.0 / deg0, deg0)
    deg1 = torch.where(deg1 > 0.0, 1.0 / deg1, deg1)
    L = torch.sparse_coo_tensor(idx, torch.cat([deg0, deg1]), (V, V))
    d = torch.arange(V, device=dev)
    L = L - torch.sparse_coo_tensor(torch.stack([d, d], dim=0), torch.ones(V, dtype=torch.float32, device=dev), (V, V))
    return L.coalesce()


def parse_weights(weight_dict, region_label_map):
    weights = np.ones(len(weight_dict))
    for region, label in region_label_map.items():
        weights[label] = weight_dict[region]
    return weights


class RegionLaplacian:
    """reset_laplacians + forward + forward_hands of RegionLaplacianLoss_v2, laplacian_type='standard'."""

    def __init__(self, verts, edges, vertex_labels, weights, sparse=False):
        lap_fn = laplacian_sparse if sparse else laplacian
        self.unique_labels = torch.unique(vertex_labels)
        self.weights = weights
        self.vertex_labels = vertex_labels
        edge_label = vertex_labels[edges]
        self.laplacians, self.vertex_partitions = [], []
        for label in self.unique_labels:
            verts_idx_included = vertex_labels == label
            selected_verts = verts[verts_idx_included]
            selected_edges = edges[torch.all(edge_label == label, dim=1)]
            _, inverse_indices = torch.unique(selected_edges, return_inverse=True)
            part_edge_local = inverse_indices.reshape(selected_edges.shape)
            self.laplacians.append(lap_fn(selected_verts, part_edge_local))
            self.vertex_partitions.append(verts_idx_included)

    def forward(self, x):
        loss = 0.0
        for i in self.unique_labels:
            x_part = x[self.vertex_partitions[i]]
            x_part = torch.matmul(self.laplacians[i].to(x.dtype), x_part)
            loss = loss + self.weights[i] * x_part.pow(2).mean()
        return loss

    def forward_hands(self, x, hand_strength=1000):
        loss = 0.0
        for i in [6, 7]:
            x_part = x[self.vertex_partitions[i]]
            x_part = torch.matmul(self.laplacians[i].to(x.dtype), x_part)
            loss = loss + hand_strength * x_part.pow(2).mean()
        return loss


def build_edges(verts, K=9):
    """loss_items.py:194-202 with knn_points restated by brute force (oracle/knn_oracle.py)."""
    from oracle.knn_oracle import knn_points
    _, idx, _ = knn_points(verts[None], verts[None], K=K + 1)
    knn_idx = idx.squeeze(0)[:, 1:]
    indices = torch.arange(verts.shape[0]).unsqueeze(1)
    return torch.cat([indices.repeat(1, K).reshape(-1, 1), knn_idx.reshape(-1, 1)], dim=1)


def pcd_laplacian_smoothing(verts, edges):
    with torch.no_grad():
        L = laplacian(verts, edges)
    loss = L.to(verts.dtype).mm(verts)
    loss = loss.norm(dim=1)
    return loss.mean()


def l2norm(human_gs_out, lambda_xyz_offsets=0.005, lambda_scales_diff=0.005, lambda_max_scale=0.001,
           max_scale_threshold=0.008, lambda_min_opacity=0.0001, min_opacity_threshold=0.2):
    xyz_offsets = human_gs_out["xyz_offsets"]
    scales = human_gs_out["scales"][:, 0]
    scales_diff = scales - scales.mean(dim=0)
    scale_thresh_idxs = scales > max_scale_threshold
    loss = lambda_xyz_offsets * xyz_offsets.norm() + lambda_scales_diff * scales_diff.norm() + \
        lambda_max_scale * scales[scale_thresh_idxs].norm()
    if "opacity" in human_gs_out:
        opacity = human_gs_out["opacity"]
        opacity_thresh_idx = opacity < min_opacity_threshold
        loss = loss + lambda_min_opacity * (0.5 - opacity[opacity_thresh_idx]).norm()
    return loss
