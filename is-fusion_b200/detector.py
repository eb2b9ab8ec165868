"""ISFusionDetector point branch (ref: mmdet3d/models/detectors/isfusion.py:13-176): the fusion-neck hot path
`extract_pts_feat` = dynamic voxelize -> DynamicVFE -> SparseEncoder -> pillarize -> ISFusionEncoder (+SECONDV2
stages) wired from the reference's config keys.  The image backbone is outside the path: its feature
maps are taken as inputs.  `pts_neck` (SECONDFPN) and `pts_bbox_head` (TransFusionHeadV2, inference forward + box
decoding: `simple_test_pts`) are the rows after the path and are built when the config carries them."""
import torch
import torch.nn.functional as F
from torch import nn

from . import fusion_encoder, head, img_backbone, img_neck, sparse_encoder, voxel_encoder  # noqa: F401  (register the modules)
from .registry import (DETECTORS, MODELS, NECKS, build_backbone, build_fusion_layer, build_middle_encoder,
                       build_voxel_encoder)
from .profiling import stage
from .voxel_layer import Voxelization


@DETECTORS.register_module()
class ISFusionDetector(nn.Module):

    def __init__(self, pts_voxel_layer=None, pts_voxel_encoder=None, pts_middle_encoder=None, fusion_encoder=None,
                 pts_backbone=None, pts_neck=None, pts_bbox_head=None, img_backbone=None, img_neck=None,
                 train_cfg=None, test_cfg=None, pretrained=None, init_cfg=None, **kwargs):
        super().__init__()
        self.pts_voxel_layer = Voxelization(**pts_voxel_layer)
        self.pts_voxel_encoder = build_voxel_encoder(dict(pts_voxel_encoder))
        self.pts_middle_encoder = build_middle_encoder(dict(pts_middle_encoder))
        self.fusion_encoder = build_fusion_layer(dict(fusion_encoder))
        self.pts_backbone = build_backbone(dict(pts_backbone))
        # pts_neck (SECONDFPN) is the next row after the hot path; built when the config carries it
        self.pts_neck = NECKS.build(dict(pts_neck)) if pts_neck is not None else None
        # img_neck (GeneralizedLSSFPN) turns the image backbone's stage outputs into img_feats; the
        # Swin backbone itself stays the reference's module
        self.img_neck = NECKS.build(dict(img_neck)) if img_neck is not None else None
        self.img_backbone = MODELS.build(dict(img_backbone)) if img_backbone is not None else None
        # pts_bbox_head (TransFusionHeadV2): like MVXTwoStageDetector, train_cfg.pts / test_cfg.pts are handed to the head
        self.pts_bbox_head = None
        if pts_bbox_head is not None:
            hc = dict(pts_bbox_head)
            hc.setdefault('train_cfg', train_cfg.get('pts') if train_cfg else None)
            hc.setdefault('test_cfg', test_cfg.get('pts') if test_cfg else None)
            self.pts_bbox_head = MODELS.build(hc)
        self.detach = kwargs.get('detach', False)
        out_size_factor = kwargs.get('out_size_factor', None)
        voxel_size = kwargs.get('voxel_size', None)
        self.pc_range = kwargs.get('pc_range', None)
        self.pillar_size = [voxel_size[0] * out_size_factor, voxel_size[1] * out_size_factor,
                            self.pc_range[5] - self.pc_range[2]]
        self.pts_pillar_layer = Voxelization(max_num_points=self.fusion_encoder.num_points_in_pillar,
                                             voxel_size=self.pillar_size, max_voxels=(30000, 60000),
                                             point_cloud_range=self.pc_range)

    @torch.no_grad()
    def dynamic_voxelize(self, points):
        """ref: isfusion.py:123-146."""
        coors = [F.pad(self.pts_voxel_layer(res), (1, 0), mode='constant', value=i) for i, res in enumerate(points)]
        return torch.cat(points, dim=0), torch.cat(coors, dim=0)

    @torch.no_grad()
    def voxelize(self, points, voxel_type='pillar'):
        """ref: isfusion.py:148-176."""
        voxels, coors, num_points = [], [], []
        for i, res in enumerate(points):
            v, c, n = self.pts_pillar_layer(res)
            voxels.append(v)
            coors.append(F.pad(c, (1, 0), mode='constant', value=i))
            num_points.append(n)
        return torch.cat(voxels, dim=0), torch.cat(num_points, dim=0), torch.cat(coors, dim=0)

    def isfusion(self, pts, pts_feats, img_feats, img_metas, batch_size, nosync=False, **kwargs):
        """ref: isfusion.py:84-101."""
        with stage('a5 pillarize'):
            if 'pts_metas' in kwargs:          # pillar branch already run on a side stream
                pts_metas = kwargs.pop('pts_metas')
            elif nosync:
                pts_metas = dict(pillars_list=self.voxelize_nosync(pts), pts=pts, pillar_size=self.pillar_size)
            else:
                pillars, pillars_num_points, pillar_coors = self.voxelize(pts, voxel_type='pillar')
                pts_metas = dict(pillars=pillars, pillars_num_points=pillars_num_points, pillar_coors=pillar_coors,
                                 pts=pts, pillar_size=self.pillar_size)
        kwargs.update(dict(pts_metas=pts_metas, img_metas=img_metas, pts_backbone=self.pts_backbone))
        return self.fusion_encoder(img_feats, pts_feats, batch_size, **kwargs)

    @torch.no_grad()
    def voxelize_nosync(self, points):
        """pillars per sample in capacity buffers with device-side counts (no host read)."""
        from .voxel_layer import hard_voxelize
        layer = self.pts_pillar_layer
        cells = int(layer.grid_size[0] * layer.grid_size[1] * layer.grid_size[2])
        max_voxels = layer.max_voxels[0] if self.training else layer.max_voxels[1]
        cap = min(cells, max_voxels)
        out = []
        for res in points:
            voxels = torch.empty((cap, layer.max_num_points, res.size(1)), dtype=res.dtype, device=res.device)
            coors = torch.empty((cap, 3), dtype=torch.int32, device=res.device)
            npv = torch.empty((cap,), dtype=torch.int32, device=res.device)
            cnt = hard_voxelize(res, voxels, coors, npv, layer.voxel_size, layer.point_cloud_range, layer.max_num_points,
                                cap, 3, True, nosync=True)
            out.append((voxels, coors, cnt))
        return out

    def _streams(self, device):
        st = self.__dict__.get('_side_streams')
        if st is None:
            st = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
            self.__dict__['_side_streams'] = st
        return st

    @torch.no_grad()
    def extract_pts_feat(self, pts, img_feats, img_metas, nosync=False, overlap=False, **kwargs):
        """ref: isfusion.py:103-121 without pts_neck; returns ([feat_0, feat_1], ins_heatmap).
        nosync=True keeps every data-dependent count on the device (capacity buffers): no host synchronisation
        anywhere in the path, which makes it CUDA-graph capturable (see GraphedNeck)."""
        batch_size = len(pts)
        rb_stream = None
        if overlap and nosync:
            # independent branches on side streams (captured as parallel branches of the CUDA graph):
            #   pillar stream : pillarize -> camera->BEV      rulebook stream : all sparse-conv index work
            main = torch.cuda.current_stream()
            pillar_stream, rb_stream = self._streams(pts[0].device)
            pillar_stream.wait_stream(main)
            with torch.cuda.stream(pillar_stream):
                pm = dict(pillars_list=self.voxelize_nosync(pts), pts=pts, pillar_size=self.pillar_size)
                enc = self.fusion_encoder
                cat = enc.new_fusion_buffer(batch_size, pts[0].device)
                enc.img_fv_to_bev([img_feats[1]], batch_size, out=cat, pts_metas=pm, img_metas=img_metas, **kwargs)
            kwargs['pts_metas'] = pm
            kwargs['img_bev_buffer'] = cat
        with stage('a1 dynamic_voxelize'):
            voxels, coors = self.dynamic_voxelize(pts)
        if overlap and nosync:
            # coordinates first: the rulebook stream starts on them while the VFE feature kernel runs
            vfe, enc3d = self.pts_voxel_encoder, self.pts_middle_encoder
            with stage('a3 dynamic_vfe'):
                vcoors, nv, ws = vfe.group(voxels, coors, batch_size)
                pre = enc3d.start_rulebooks(vcoors, batch_size, nv, rb_stream)
                vfeats = vfe.compute(voxels, batch_size, ws)
            with stage('a4 sparse_encoder'):
                x, _, kwargs = enc3d(vfeats, vcoors, batch_size, dense=False, prebuilt=pre, **kwargs)
        else:
            with stage('a3 dynamic_vfe'):
                r = self.pts_voxel_encoder(voxels, coors, pts, img_feats, img_metas, batch_size=batch_size, nosync=nosync)
            nv = r[2] if nosync else None
            with stage('a4 sparse_encoder'):
                x, _, kwargs = self.pts_middle_encoder(r[0], r[1], batch_size, dense=False, num_voxels_dev=nv,
                                                       sorted_coors=True, **kwargs)
        if overlap and nosync:
            torch.cuda.current_stream().wait_stream(self._streams(pts[0].device)[0])
        kwargs['lidar_sparse'] = self.pts_middle_encoder.last_out
        feats, ins_hm = self.isfusion(pts, x, img_feats, img_metas, batch_size, nosync=nosync, **kwargs)
        if self.pts_neck is not None:        # isfusion.py:115-116
            feats = self.pts_neck(feats)
        return feats, ins_hm

    def forward(self, *args, return_loss=None, **kwargs):
        """`net(points, img_feats, img_metas, ...)` runs the path itself (extract_pts_feat).  With the reference's keyword
        (detectors/base.py:46-62) `return_loss=False` dispatches to forward_test (double-nested points / img_metas / img);
        `return_loss=True` is the training step, which is outside this build."""
        if return_loss is None:
            return self.extract_pts_feat(*args, **kwargs)
        if return_loss:
            raise NotImplementedError('ISFusionDetector: forward_train (losses, assigner) is outside the path')
        return self.forward_test(*args, **kwargs)

    def extract_feat(self, points, img, img_metas, **kwargs):
        """ref: isfusion.py:178-182."""
        img_feats = self.extract_img_feat(img, img_metas)
        pts_feats = self.extract_pts_feat(points, img_feats, img_metas, **kwargs)
        return (img_feats, pts_feats)

    def forward_test(self, points, img_metas=None, img=None, **kwargs):
        """ref: detectors/base.py:15-44: the outer lists are test-time augmentations; one entry -> simple_test."""
        for var, name in [(points, 'points'), (img_metas, 'img_metas')]:
            if not isinstance(var, list):
                raise TypeError('{} must be a list, but got {}'.format(name, type(var)))
        if len(points) != len(img_metas):
            raise ValueError('num of augmentations ({}) != num of image meta ({})'.format(len(points), len(img_metas)))
        if len(points) != 1:
            raise NotImplementedError('ISFusionDetector: aug_test (test-time augmentation) is outside the path')
        img = [img] if img is None else img
        return self.simple_test(points[0], img_metas[0], img[0], **kwargs)

    @torch.no_grad()
    def extract_img_feat(self, img, img_metas):
        """ref: isfusion.py:54-82: img [B, N, 3, H, W] -> img_feats (tuple of [B*N, 256, h, w]); records input_shape."""
        if self.img_backbone is None or img is None:
            return None
        input_shape = tuple(img.shape[-2:])
        for meta in img_metas:
            meta.update(input_shape=input_shape)
        b, n, c, h, w = img.shape
        feats = self.img_backbone(img.view(b * n, c, h, w))
        return self.img_neck(feats) if self.img_neck is not None else feats

    @torch.no_grad()
    def simple_test_pts(self, x, x_img=None, img_metas=None, rescale=False):
        """ref: isfusion.py:272-281: head forward + box decoding on the pts_neck output.  Returns per sample
        dict(boxes_3d [K,9] tensor, scores_3d, labels_3d) — `bbox3d2result` without the mmdet3d box class."""
        if self.pts_bbox_head is None:
            raise RuntimeError('ISFusionDetector: the config carries no pts_bbox_head')
        outs = self.pts_bbox_head(x, x_img, img_metas)
        return [dict(boxes_3d=b['bboxes'], scores_3d=b['scores'], labels_3d=b['labels'])
                for b in self.pts_bbox_head.get_bboxes(outs, img_metas, rescale=rescale)]

    @torch.no_grad()
    def simple_test(self, points, img_metas, img=None, img_feats=None, rescale=False, **kwargs):
        """ref: isfusion.py:283-302; `img` [B, N, 3, H, W] runs the image branch, or pass its output as `img_feats`."""
        if img_feats is None:
            img_feats = self.extract_img_feat(img, img_metas)
        if self.pts_neck is None:
            raise RuntimeError('ISFusionDetector.simple_test: the head consumes the pts_neck output (config:105-112)')
        feats, _ = self.extract_pts_feat(points, img_feats, img_metas, **kwargs)
        return [dict(pts_bbox=r) for r in self.simple_test_pts(feats, img_feats, img_metas, rescale=rescale)]


class GraphedNeck:
    """The whole fusion-neck forward as ONE CUDA graph replay.

    Capacity buffers + device-side counts make every launch shape static, so the ~400 kernels of a frame are
    captured once and replayed without any host work between them.  Frames with fewer points than `num_points`
    are padded with out-of-range points (dropped by voxelization).  `capacities` = output-site buffer rows of
    the five strided sparse convs (calibrate() measures them on sample frames); check() reports overflow.
    """

    def __init__(self, net, batch, num_points, img_hw=(256, 704), feat_hw=None, calib=None, capacities=None,
                 device=None, overlap=True):
        import isfusion_b200.spconv as spmod
        self.net, self.batch, self.num_points = net, batch, num_points
        self.overlap = overlap
        dev = device or next(net.parameters()).device
        self.dev = dev
        fh, fw = feat_hw or (img_hw[0] // 16, img_hw[1] // 16)
        self.points = [torch.full((num_points, 5), 1e6, dtype=torch.float32, device=dev) for _ in range(batch)]
        self.img_feat = torch.zeros((batch * net.fusion_encoder.num_views, net.fusion_encoder.embed_dims, fh, fw),
                                    dtype=torch.float32, device=dev)
        self.metas = [dict(input_shape=img_hw)] * batch
        self.cam_proj = net.fusion_encoder.camera_projection(calib['lidar2img'], calib['img_aug_matrix'],
                                                              calib['lidar_aug_matrix'], batch, dev)
        if capacities is not None:
            net.pts_middle_encoder.set_capacities(capacities)
        self._spmod = spmod
        self.graph, self.outputs, self.watch = None, None, []

    def _run(self):
        feats = [self.img_feat[:0], self.img_feat, self.img_feat[:0]]
        return self.net.extract_pts_feat(self.points, feats, self.metas, nosync=True, overlap=self.overlap,
                                         cam_proj=self.cam_proj)

    def load(self, points, img_feat1):
        for dst, src in zip(self.points, points):
            n = src.shape[0]
            if n > self.num_points:
                raise ValueError('frame has %d points, graph capacity is %d' % (n, self.num_points))
            dst[:n].copy_(src, non_blocking=True)
            if n < self.num_points:
                dst[n:].fill_(1e6)
        self.img_feat.copy_(img_feat1, non_blocking=True)

    def capture(self, warmup=3):
        """warm up and capture on streams owned by THIS graph (main + the two branch streams), so that several
        GraphedNeck lanes over one set of weights never share a scratch workspace (`_lib.workspace` is keyed by
        stream) and can be replayed concurrently (FramePipeline)."""
        if self.__dict__.get('_lane') is None:
            self._lane = (torch.cuda.Stream(device=self.dev), (torch.cuda.Stream(device=self.dev),
                                                                torch.cuda.Stream(device=self.dev)))
        s, sides = self._lane
        saved = self.net.__dict__.get('_side_streams')
        self.net.__dict__['_side_streams'] = sides
        try:
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(warmup):
                    self._spmod.OVERFLOW_WATCH.clear()
                    self._run()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self._spmod.OVERFLOW_WATCH.clear()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=s):
                self.outputs = self._run()
        finally:
            self.net.__dict__['_side_streams'] = saved
        self.watch = list(self._spmod.OVERFLOW_WATCH)
        self._spmod.OVERFLOW_WATCH.clear()
        # the graph bakes the device pointers of every packed / folded weight tensor (and TMA descriptors built from them):
        # keep those cache objects alive for the life of the graph, so that a later net.eval() / load_state_dict(), which drops
        # the module-side caches, cannot hand their memory back to the allocator under a replay
        self._keepalive = [m.__dict__[k] for m in self.net.modules() for k in ('_pk', '_packed') if m.__dict__.get(k) is not None]
        return self

    def __call__(self, points=None, img_feat1=None):
        if points is not None:
            self.load(points, img_feat1)
        self.graph.replay()
        return self.outputs

    # ---- pipelined host interface: H2D of frame i+1 and D2H of frame i overlap the replay of frame i / i+1 --------
    def _pipe(self):
        p = self.__dict__.get('_pipe_state')
        if p is None:
            dev = self.dev
            p = dict(copy=torch.cuda.Stream(device=dev),
                     pts=[[torch.empty_like(t) for t in self.points] for _ in range(2)],
                     img=[torch.empty_like(self.img_feat) for _ in range(2)],
                     n=[[0] * self.batch for _ in range(2)], ev_in=[torch.cuda.Event(), torch.cuda.Event()],
                     ev_used=[None, None],
                     out=None, out_host=None, ev_out=torch.cuda.Event(), ev_d2h=torch.cuda.Event(), slot=0, pending=None)
            self.__dict__['_pipe_state'] = p
        return p

    def prefetch(self, points_host, img_feat1_host):
        """start the H2D copy of the NEXT frame (pinned host tensors) on the copy stream; returns a ticket."""
        p = self._pipe()
        k = p['slot']
        p['slot'] ^= 1
        with torch.cuda.stream(p['copy']):
            if p['ev_used'][k] is not None:
                p['copy'].wait_event(p['ev_used'][k])     # the replay that read this staging set has consumed it
            for b, src in enumerate(points_host):
                n = src.shape[0]
                if n > self.num_points:
                    raise ValueError('frame has %d points, graph capacity is %d' % (n, self.num_points))
                p['pts'][k][b][:n].copy_(src, non_blocking=True)
                p['n'][k][b] = n
            p['img'][k].copy_(img_feat1_host, non_blocking=True)
            p['ev_in'][k].record()
        return k

    def run_prefetched(self, ticket, out_host=None):
        """replay on the prefetched frame; the two output maps are copied to `out_host` (pinned) on the copy stream,
        overlapping the next replay.  Returns the device outputs (valid until the next replay)."""
        p = self._pipe()
        main = torch.cuda.current_stream()
        main.wait_event(p['ev_in'][ticket])
        for b, dst in enumerate(self.points):           # staging -> the graph's static inputs (device-to-device)
            n = p['n'][ticket][b]
            dst[:n].copy_(p['pts'][ticket][b][:n], non_blocking=True)
            if n < self.num_points:
                dst[n:].fill_(1e6)
        self.img_feat.copy_(p['img'][ticket], non_blocking=True)
        if p['ev_used'][ticket] is None:
            p['ev_used'][ticket] = torch.cuda.Event()
        p['ev_used'][ticket].record()
        self.graph.replay()
        feats, _ = self.outputs
        if out_host is not None:
            if p['out'] is None:
                p['out'] = [torch.empty_like(t) for t in feats]
            else:
                main.wait_event(p['ev_d2h'])            # the previous D2H has left the staging outputs
            for d, t in zip(p['out'], feats):
                d.copy_(t, non_blocking=True)
            p['ev_out'].record()
            with torch.cuda.stream(p['copy']):
                p['copy'].wait_event(p['ev_out'])
                for h, d in zip(out_host, p['out']):
                    h.copy_(d, non_blocking=True)
                p['ev_d2h'].record()
        return self.outputs

    def check(self):
        """[(count, capacity)] of strided convs whose output buffer overflowed in the last replay (host sync)."""
        return [(int(c.item()), cap) for c, cap in self.watch if int(c.item()) >= cap]

    @staticmethod
    def calibrate(net, frames, img_feats, metas, margin=1.3, **kw):
        """run sample frames in exact mode and size the strided-conv buffers to margin x the largest seen."""
        best = None
        for pts in frames:
            net.extract_pts_feat(pts, img_feats, metas, **kw)
            enc = net.pts_middle_encoder
            rows = []
            x = enc.last_encode_features
            rows = [f.features.shape[0] for f in x[1:4]] + [enc.last_out.features.shape[0]]
            best = rows if best is None else [max(a, b) for a, b in zip(best, rows)]
        return [int(r * margin) // 128 * 128 + 128 for r in best]


class FramePipeline:
    """`lanes` GraphedNeck graphs over ONE set of weights, replayed round-robin on their own streams: consecutive
    frames are in flight together, so the launch-bound tail of frame i (the 200-query decoder, small transformer
    GEMMs, index kernels that fill a fraction of the 148 SMs) overlaps the sparse encoder of frame i+1.  Each lane
    owns its static inputs, activations (graph-private pool), scratch workspaces and copy stream; the outputs of a
    frame stay valid until the same lane is submitted again (`lanes` - 1 frames later)."""

    def __init__(self, net, lanes, *args, **kwargs):
        self.lanes = [GraphedNeck(net, *args, **kwargs) for _ in range(lanes)]
        self.streams = [torch.cuda.Stream(device=self.lanes[0].dev) for _ in range(lanes)]
        self.turn = 0

    def capture(self, points, img_feat1, warmup=3):
        for g in self.lanes:
            g.load(points, img_feat1)
            g.capture(warmup)
        return self

    def _next(self):
        k = self.turn
        self.turn = (k + 1) % len(self.lanes)
        return k

    def submit(self, points, img_feat1):
        """resident (device) frame -> (lane, outputs); asynchronous, ordered after the caller's current stream."""
        k = self._next()
        st = self.streams[k]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            out = self.lanes[k](points, img_feat1)
        return k, out

    def prefetch(self, lane, points_host, img_feat1_host):
        return self.lanes[lane].prefetch(points_host, img_feat1_host)

    def submit_prefetched(self, lane, ticket, out_host=None):
        st = self.streams[lane]
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            return self.lanes[lane].run_prefetched(ticket, out_host)

    def join(self):
        """the caller's current stream waits for every frame in flight (replays and their D2H copies)."""
        cur = torch.cuda.current_stream()
        for g, st in zip(self.lanes, self.streams):
            cur.wait_stream(st)
            p = g.__dict__.get('_pipe_state')
            if p is not None:
                cur.wait_stream(p['copy'])

    def check(self):
        return [o for g in self.lanes for o in g.check()]
