"""Oracle: voxelization, dynamic scatter, DynamicVFE (rows a1, a2, a3, a5 of SURVEY §8).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def grid_size_of(voxel_size, coors_range):
    """ref: mmdet3d/ops/voxel/src/voxelization_cuda.cu:508-510 /
    voxelization_cpu.cpp:124-127  grid = round((max-min)/v) with float32 operands."""
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    return [int(round(float((cr[3 + i] - cr[i]) / vs[i]))) for i in range(3)]


def dynamic_voxelize(points, voxel_size, coors_range, mode='cuda'):
    """points [N,>=3] fp32 -> coors [N,3] int32 in (z,y,x) order.

    ref: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:7-43 (mode='cpu': a failed
    point gets -1 in all three slots) and voxelization_cuda.cu:24-61 (mode='cuda':
    only the slots up to the first failing axis are written with -1, the rest keep
    the caller's zero initialisation, voxelize.py:53).  Arithmetic is fp32
    subtract -> fp32 divide -> floor, exactly as both sources do.
    """
    p = np.ascontiguousarray(points.detach().cpu().numpy() if torch.is_tensor(points) else points,
                             dtype=np.float32)
    vs = np.asarray(voxel_size, dtype=np.float32)
    cr = np.asarray(coors_range, dtype=np.float32)
    grid = grid_size_of(voxel_size, coors_range)
    n = p.shape[0]
    c = np.empty((n, 3), dtype=np.int64)
    for j in range(3):
        q = (p[:, j] - cr[j]) / vs[j]          # fp32 sub then fp32 div (numpy keeps float32)
        assert q.dtype == np.float32
        c[:, j] = np.floor(q).astype(np.int64)
    bad = [(c[:, j] < 0) | (c[:, j] >= grid[j]) for j in range(3)]
    out = np.zeros((n, 3), dtype=np.int32)
    if mode == 'cpu':
        fail = bad[0] | bad[1] | bad[2]
        out[:, 0] = c[:, 2]
        out[:, 1] = c[:, 1]
        out[:, 2] = c[:, 0]
        out[fail] = -1
    else:
        fx = bad[0]
        fy = ~fx & bad[1]
        fz = ~fx & ~fy & bad[2]
        ok = ~(fx | fy | fz)
        out[ok, 0] = c[ok, 2]
        out[ok, 1] = c[ok, 1]
        out[ok, 2] = c[ok, 0]
        out[fx, 0] = -1
        out[fy, 0] = -1
        out[fy, 1] = -1
        out[fz] = -1
    return torch.from_numpy(out)


def hard_voxelize_loop(points, voxel_size, coors_range, max_points, max_voxels):
    """Literal loop restatement (small inputs only).
    ref: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:45-101."""
    pts = points.detach().cpu().numpy().astype(np.float32)
    coor = dynamic_voxelize(points, voxel_size, coors_range, mode='cpu').numpy()
    gx, gy, gz = grid_size_of(voxel_size, coors_range)
    c2v = -np.ones((gz, gy, gx), dtype=np.int64)
    nf = pts.shape[1]
    voxels = np.zeros((max_voxels, max_points, nf), dtype=np.float32)
    coors = np.zeros((max_voxels, 3), dtype=np.int32)
    npv = np.zeros((max_voxels,), dtype=np.int32)
    voxel_num = 0
    for i in range(pts.shape[0]):
        if coor[i, 0] == -1:
            continue
        vi = c2v[coor[i, 0], coor[i, 1], coor[i, 2]]
        if vi == -1:
            vi = voxel_num
            if max_voxels != -1 and voxel_num >= max_voxels:
                continue
            voxel_num += 1
            c2v[coor[i, 0], coor[i, 1], coor[i, 2]] = vi
            coors[vi] = coor[i]
        num = npv[vi]
        if max_points == -1 or num < max_points:
            voxels[vi, num] = pts[i]
            npv[vi] += 1
    return (torch.from_numpy(voxels[:voxel_num]), torch.from_numpy(coors[:voxel_num]),
            torch.from_numpy(npv[:voxel_num]))


def hard_voxelize(points, voxel_size, coors_range, max_points, max_voxels):
    """Vectorised restatement of the same first-come semantics (any N).

    ref: mmdet3d/ops/voxel/src/voxelization_cpu.cpp:45-101 (== the deterministic
    CUDA path voxelization_cuda.cu:231-373): voxels are numbered in order of the
    first point (by point index) that falls in them; a voxel keeps its first
    ``max_points`` points by index; voxels numbered >= max_voxels are dropped.
    Checked against hard_voxelize_loop and oracle/_ref in tests.
    """
    pts = points.detach().cpu().numpy().astype(np.float32)
    coor = dynamic_voxelize(points, voxel_size, coors_range, mode='cpu').numpy().astype(np.int64)
    gx, gy, gz = grid_size_of(voxel_size, coors_range)
    valid = coor[:, 0] >= 0
    idx = np.nonzero(valid)[0]
    key = (coor[idx, 0] * gy + coor[idx, 1]) * gx + coor[idx, 2]
    order = np.argsort(key, kind='stable')            # groups by voxel, index order kept inside
    skey = key[order]
    sidx = idx[order]
    nf = pts.shape[1]
    if len(skey) == 0:
        return (torch.zeros((0, max_points, nf)), torch.zeros((0, 3), dtype=torch.int32),
                torch.zeros((0,), dtype=torch.int32))
    head = np.ones(len(skey), dtype=bool)
    head[1:] = skey[1:] != skey[:-1]
    seg_start = np.nonzero(head)[0]
    seg_id = np.cumsum(head) - 1
    rank_in = np.arange(len(skey)) - seg_start[seg_id]
    first_pt = sidx[seg_start]                          # first point index of each voxel
    vorder = np.argsort(first_pt, kind='stable')        # voxel numbering = order of first appearance
    vnum = np.empty_like(vorder)
    vnum[vorder] = np.arange(len(vorder))
    M = len(seg_start) if max_voxels == -1 else min(len(seg_start), max_voxels)
    counts = np.diff(np.append(seg_start, len(skey)))
    voxels = np.zeros((M, max_points, nf), dtype=np.float32)
    coors = np.zeros((M, 3), dtype=np.int32)
    npv = np.zeros((M,), dtype=np.int32)
    pv = vnum[seg_id]
    keep = (pv < M) & (rank_in < max_points)
    voxels[pv[keep], rank_in[keep]] = pts[sidx[keep]]
    keepv = vnum < M
    coors[vnum[keepv]] = coor[first_pt[keepv]].astype(np.int32)
    npv[vnum[keepv]] = np.minimum(counts[keepv], max_points).astype(np.int32)
    return torch.from_numpy(voxels), torch.from_numpy(coors), torch.from_numpy(npv)


def dynamic_scatter(feats, coors, reduce_type='max'):
    """feats [N,C] fp32, coors [N,ndim] int -> (reduced [V,C], out_coors [V,ndim],
    coors_map [N] int32 (-1 for dropped rows), reduce_count [V] int32).

    ref: mmdet3d/ops/voxel/src/scatter_points_cuda.cu:183-239: rows with any
    negative coordinate are dropped (:202); out_coors = at::unique_dim(sorted,
    lexicographic) (:204-205); 'mean' = sum / count (:233-234); 'max' starts from
    -inf (:220).  Brute-force definition: tests/.../test_dynamic_scatter.py:57-66.
    """
    feats = feats.detach().cpu().float()
    coors = coors.detach().cpu()
    n, c = feats.shape
    if n == 0:
        return (feats.clone(), coors.clone(), torch.zeros(0, dtype=torch.int32),
                torch.zeros(0, dtype=torch.int32))
    valid = (coors >= 0).all(dim=1)
    vc = coors[valid]
    out_coors, inv, counts = torch.unique(vc, dim=0, sorted=True, return_inverse=True,
                                          return_counts=True)
    v = out_coors.shape[0]
    cmap = torch.full((n,), -1, dtype=torch.int64)
    cmap[valid] = inv
    vf = feats[valid]
    if reduce_type in ('sum', 'mean'):
        red = torch.zeros((v, c), dtype=torch.float32).index_add_(0, inv, vf)
        if reduce_type == 'mean':
            red = red / counts.unsqueeze(-1).to(red.dtype)
    else:
        red = torch.full((v, c), float('-inf'))
        red = red.scatter_reduce(0, inv.unsqueeze(1).expand(-1, c), vf, reduce='amax', include_self=True)
    return red, out_coors.to(coors.dtype), cmap.to(torch.int32), counts.to(torch.int32)


def dynamic_scatter_backward(grad_reduced, feats, reduced, coors_map, reduce_count, reduce_type='max'):
    """gradient of dynamic_scatter w.r.t. feats, as the reference's backward kernels define it
    (scatter_points_cuda.cu:105-179, driver :241-320; NO CPU path in the reference, voxelization.h:139, so this is a
    literal loop restatement — for sum / mean it equals autograd of the brute-force definition, which the test checks):
      sum : grad_feats[x] = grad_reduced[map[x]]   (:121-124)      mean: ... / reduce_count[map[x]]   (:125-130)
      max : per (voxel, channel) the LOWEST point index with feats == reduced gets the gradient (atomicMin, :135-161;
            scatter :163-179); everything else, and rows with map == -1, stay 0 (grad_feats.fill_(0), :258)."""
    import numpy as np
    g = grad_reduced.detach().cpu().float().numpy()
    f = feats.detach().cpu().float().numpy()
    r = reduced.detach().cpu().float().numpy()
    m = coors_map.detach().cpu().numpy()
    cnt = reduce_count.detach().cpu().numpy()
    n, c = f.shape
    out = np.zeros((n, c), np.float32)
    if reduce_type in ('sum', 'mean'):
        for x in range(n):
            if m[x] < 0:
                continue
            out[x] = g[m[x]] if reduce_type == 'sum' else g[m[x]] / np.float32(cnt[m[x]])
        return torch.from_numpy(out)
    reduce_from = np.full(r.shape, n, np.int64)
    for x in range(n):
        if m[x] < 0:
            continue
        for i in range(c):
            if f[x, i] == r[m[x], i] and x < reduce_from[m[x], i]:
                reduce_from[m[x], i] = x
    for v in range(r.shape[0]):
        for i in range(c):
            if reduce_from[v, i] < n:
                out[reduce_from[v, i], i] = g[v, i]
    return torch.from_numpy(out)


class DynamicScatter(nn.Module):
    """ref: mmdet3d/ops/voxel/scatter_points.py:54-103 (per-sample loop + batch pad)."""

    def __init__(self, voxel_size, point_cloud_range, average_points):
        super().__init__()
        self.average_points = average_points

    def forward(self, points, coors):
        reduce = 'mean' if self.average_points else 'max'
        if coors.size(-1) == 3:
            r = dynamic_scatter(points, coors, reduce)
            return r[0], r[1]
        batch_size = int(coors[-1, 0]) + 1
        voxels, voxel_coors = [], []
        for i in range(batch_size):
            inds = torch.where(coors[:, 0] == i)
            v, vc, _, _ = dynamic_scatter(points[inds], coors[inds][:, 1:], reduce)
            voxel_coors.append(F.pad(vc, (1, 0), mode='constant', value=i))
            voxels.append(v)
        return torch.cat(voxels, 0), torch.cat(voxel_coors, 0)


class DynamicVFELayer(nn.Module):
    """ref: mmdet3d/models/voxel_encoders/utils.py:107-144 (Linear(no bias) -> BN1d -> ReLU)."""

    def __init__(self, in_channels, out_channels, eps=1e-3, momentum=0.01):
        super().__init__()
        self.norm = nn.BatchNorm1d(out_channels, eps=eps, momentum=momentum)
        self.linear = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, inputs):
        return F.relu(self.norm(self.linear(inputs)))


class DynamicVFE(nn.Module):
    """ref: mmdet3d/models/voxel_encoders/voxel_encoder.py:287-547 (cluster-centre +
    voxel-centre offsets, 2 VFE layers, max pooling).  state_dict keys match the
    reference: vfe_layers.{i}.{linear.weight, norm.*}."""

    def __init__(self, in_channels=4, feat_channels=(), with_distance=False, with_cluster_center=False,
                 with_voxel_center=False, voxel_size=(0.2, 0.2, 4), point_cloud_range=(0, -40, -3, 70.4, 40, 1),
                 norm_cfg=dict(type='BN1d', eps=1e-3, momentum=0.01), mode='max', **kwargs):
        super().__init__()
        assert mode == 'max' and not with_distance
        if with_cluster_center:
            in_channels += 3
        if with_voxel_center:
            in_channels += 3
        self.in_channels = in_channels
        self._with_cluster_center = with_cluster_center
        self._with_voxel_center = with_voxel_center
        self.vx, self.vy, self.vz = voxel_size
        self.x_offset = self.vx / 2 + point_cloud_range[0]
        self.y_offset = self.vy / 2 + point_cloud_range[1]
        self.z_offset = self.vz / 2 + point_cloud_range[2]
        self.point_cloud_range = point_cloud_range
        chans = [self.in_channels] + list(feat_channels)
        layers = []
        for i in range(len(chans) - 1):
            cin = chans[i] * (2 if i > 0 else 1)
            layers.append(DynamicVFELayer(cin, chans[i + 1], norm_cfg.get('eps', 1e-5), norm_cfg.get('momentum', 0.1)))
        self.vfe_layers = nn.ModuleList(layers)
        self.vfe_scatter = DynamicScatter(voxel_size, point_cloud_range, False)
        self.cluster_scatter = DynamicScatter(voxel_size, point_cloud_range, True)

    def map_voxel_center_to_point(self, pts_coors, voxel_mean, voxel_coors):
        """ref: voxel_encoder.py:410-450.  The reference scatters voxel row ids into a
        dense B*Z*Y*X int64 canvas (663 MB at the config's grid) and reads it back at
        each point's linear index.  For in-range points (Appendix C #5) that is the
        map point -> row of its voxel, which is computed here by sorted search on the
        same linear index (no canvas)."""
        cz = round((self.point_cloud_range[5] - self.point_cloud_range[2]) / self.vz)
        cy = round((self.point_cloud_range[4] - self.point_cloud_range[1]) / self.vy)
        cx = round((self.point_cloud_range[3] - self.point_cloud_range[0]) / self.vx)

        def lin(c):
            c = c.long()
            return c[:, 0] * cz * cy * cx + c[:, 1] * cy * cx + c[:, 2] * cx + c[:, 3]
        vkey = lin(voxel_coors)                  # sorted ascending (unique_dim order, batch-major)
        pkey = lin(pts_coors)
        pos = torch.searchsorted(vkey, pkey)
        assert (vkey[pos] == pkey).all()
        return voxel_mean[pos]

    @torch.no_grad()
    def forward(self, features, coors, *args, **kwargs):
        features = features.float()
        features_ls = [features]
        if self._with_cluster_center:
            voxel_mean, mean_coors = self.cluster_scatter(features, coors)
            points_mean = self.map_voxel_center_to_point(coors, voxel_mean, mean_coors)
            features_ls.append(features[:, :3] - points_mean[:, :3])
        if self._with_voxel_center:
            f_center = features.new_zeros(size=(features.size(0), 3))
            f_center[:, 0] = features[:, 0] - (coors[:, 3].type_as(features) * self.vx + self.x_offset)
            f_center[:, 1] = features[:, 1] - (coors[:, 2].type_as(features) * self.vy + self.y_offset)
            f_center[:, 2] = features[:, 2] - (coors[:, 1].type_as(features) * self.vz + self.z_offset)
            features_ls.append(f_center)
        features = torch.cat(features_ls, dim=-1)
        for i, vfe in enumerate(self.vfe_layers):
            point_feats = vfe(features)
            voxel_feats, voxel_coors = self.vfe_scatter(point_feats, coors)
            if i != len(self.vfe_layers) - 1:
                feat_per_point = self.map_voxel_center_to_point(coors, voxel_feats, voxel_coors)
                features = torch.cat([point_feats, feat_per_point], dim=1)
        return voxel_feats, voxel_coors


def dynamic_vfe_train_forward(vfe, features, coors):
    """DIFFERENTIABLE restatement of DynamicVFE.forward for training-parity tests (the forward above runs under no_grad
    because the reference's scatter has no CPU autograd): same algorithm (voxel_encoder.py:470-547) with the scatters written
    as torch index_add / scatter_reduce on the lexicographically sorted unique coordinates (= the order of `dynamic_scatter`).
    The max-pool gradient of torch (split among ties) and of the reference (lowest-index arg-max, scatter_points_cuda.cu:105-179)
    differ only on ties, which after the ReLU occur at 0 where the ReLU gradient is 0 anyway."""
    features = features.float()
    uniq, inv = torch.unique(coors, dim=0, sorted=True, return_inverse=True)
    v = uniq.shape[0]
    cnt = torch.zeros(v).index_add_(0, inv, torch.ones(features.shape[0]))
    mean = torch.zeros(v, 3).index_add_(0, inv, features[:, :3]) / cnt[:, None]
    cf = coors.float()
    f_center = torch.stack([features[:, 0] - (cf[:, 3] * vfe.vx + vfe.x_offset), features[:, 1] - (cf[:, 2] * vfe.vy + vfe.y_offset),
                            features[:, 2] - (cf[:, 1] * vfe.vz + vfe.z_offset)], 1)
    x = torch.cat([features, features[:, :3] - mean[inv], f_center], 1)
    vf = None
    for i, layer in enumerate(vfe.vfe_layers):
        pf = layer(x)
        vf = torch.full((v, pf.shape[1]), float('-inf')).scatter_reduce(0, inv[:, None].expand(-1, pf.shape[1]), pf, reduce='amax',
                                                                         include_self=True)
        if i != len(vfe.vfe_layers) - 1:
            x = torch.cat([pf, vf[inv]], 1)
    return vf, uniq.to(coors.dtype)


class Voxelization(nn.Module):
    """ref: mmdet3d/ops/voxel/voxelize.py:77-148."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        if self.max_num_points == -1 or max_voxels == -1:
            return dynamic_voxelize(input, self.voxel_size, self.point_cloud_range, mode='cuda')
        return hard_voxelize(input, self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels)
