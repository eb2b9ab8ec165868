"""Generate the golden fixtures under tests/golden/ FROM THE REFERENCE ITSELF.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
  * voxel_generator_kat.npz : the reference's numpy VoxelGenerator (mmdet3d/core/voxel/
    voxel_generator.py:212-280) on the known-answer input of
    tests/test_models/test_voxel_encoder/test_voxel_generator.py:7-23.
  * voxel_ref_cpp.npz       : the reference's C++ dynamic/hard voxelization
    (mmdet3d/ops/voxel/src/voxelization_cpu.cpp, compiled by oracle/build_ref.py) on seeded
    clouds incl. an adversarial set with coordinates within 1 ulp of cell faces.
  * msda_ref.npz            : ops/functions/ms_deform_attn_func.py:41-61
    (ms_deform_attn_core_pytorch) on the shapes of ops/test.py:20-36 and on the IGF shape.
  * spconv_ref_cpp.npz      : rulebooks (output sites + (offset, input row, output site) pairs) of the vendored
    spconv-1.x CPU generator (bevfusion-ops/spconv/include/spconv/geometry.h:144-297) for the four conv shapes
    of the SparseEncoder.
(The fixtures made by running the reference's Python modules are in make_golden_ref.py.)
"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = '/root/reference'


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def voxel_generator_kat():
    vg = _load('vg', os.path.join(REF, 'mmdet3d/core/voxel/voxel_generator.py'))
    np.random.seed(0)
    g = vg.VoxelGenerator([0.5, 0.5, 0.5], [0, -40, -3, 70.4, 40, 1], 1000)
    points = np.random.rand(1000, 4)
    voxels, coors, npv = g.generate(points)
    np.savez_compressed(os.path.join(HERE, 'voxel_generator_kat.npz'), points=points.astype(np.float32),
                        voxels=voxels.astype(np.float32), coors=coors.astype(np.int32), npv=npv.astype(np.int32))


def voxel_ref_cpp():
    from oracle import build_ref, ref_voxel
    build_ref.build()
    g = torch.Generator().manual_seed(0)
    n = 4000
    pts = torch.rand(n, 5, generator=g) * torch.tensor([112., 112., 8.4, 1., 0.5]) + torch.tensor([-56., -56., -5.2, 0., 0.])
    # adversarial: 10% of the points within 1 ulp of a voxel face
    vs = np.float32(0.075)
    k = torch.randint(0, 1440, (n // 10,), generator=g).numpy().astype(np.float32)
    face = (k * vs + np.float32(-54.0)).astype(np.float32)
    adv = np.stack([np.nextafter(face, np.float32(1e9)), face, np.nextafter(face, np.float32(-1e9))], 0)
    sel = torch.randint(0, 3, (n // 10,), generator=g).numpy()
    pts[: n // 10, 0] = torch.from_numpy(adv[sel, np.arange(n // 10)])
    pts[n // 10: n // 5, 1] = torch.from_numpy(adv[(sel + 1) % 3, np.arange(n // 10)])
    out = dict(points=pts.numpy())
    cfgs = {'main': ([0.075, 0.075, 0.2], [-54, -54, -5, 54, 54, 3]),
            'pillar': ([0.6, 0.6, 8.0], [-54, -54, -5, 54, 54, 3])}
    for name, (v, r) in cfgs.items():
        out['dyn_' + name] = ref_voxel.dynamic_voxelize(pts, v, r).numpy()
    for tag, mp, mv in [('p12_v60000', 12, 60000), ('p12_v300', 12, 300), ('p3_v60000', 3, 60000)]:
        vo, co, nu = ref_voxel.hard_voxelize(pts, *cfgs['pillar'], mp, mv)
        out['hard_%s_voxels' % tag] = vo.numpy()
        out['hard_%s_coors' % tag] = co.numpy()
        out['hard_%s_npv' % tag] = nu.numpy()
    np.savez_compressed(os.path.join(HERE, 'voxel_ref_cpp.npz'), **out)


def msda_ref():
    f = _load('msda', os.path.join(REF, 'ops/functions/ms_deform_attn_func.py')) if False else None
    # ms_deform_attn_func.py imports the compiled MultiScaleDeformableAttention at module import;
    # exec only the pure-PyTorch function (lines 41-61) from the reference file, unmodified.
    src = open(os.path.join(REF, 'ops/functions/ms_deform_attn_func.py')).read().split('\n')
    body = '\n'.join(src[40:61])
    ns = {'torch': torch, 'F': torch.nn.functional}
    exec(body, ns)
    core = ns['ms_deform_attn_core_pytorch']
    out = {}
    torch.manual_seed(3)
    # case A: ops/test.py:20-36 shapes
    N, M, D, Lq, Lv, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, Lv, P, 2)
    aw = torch.rand(N, Lq, M, Lv, P) + 1e-5
    aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out.update(a_value=value.numpy(), a_shapes=shapes.numpy(), a_loc=loc.numpy(), a_aw=aw.numpy(),
               a_out=core(value.double(), shapes, loc.double(), aw.double()).float().numpy())
    # case B: the IGF head layout (fusion_encoder.py:773-780): 1 level, 8 heads x 16 ch, 16 points, 200 queries
    N, M, D, Lq, Lv, P = 1, 8, 16, 200, 1, 16
    shapes = torch.as_tensor([(40, 44)], dtype=torch.long)   # small non-square map, IGF head layout
    value = torch.randn(N, 40 * 44, M, D).half().float()
    loc = torch.rand(N, Lq, M, Lv, P, 2) * 1.2 - 0.1          # some samples fall outside [0,1]
    aw = torch.softmax(torch.randn(N, Lq, M, Lv * P), -1).view(N, Lq, M, Lv, P)
    ref = core(value.double(), shapes, loc.double(), aw.double()).float()
    out.update(b_seed=np.int64(3), b_loc=loc.numpy().astype(np.float32), b_aw=aw.numpy().astype(np.float32),
               b_value_f16=value.half().numpy(), b_out=ref.numpy())
    np.savez_compressed(os.path.join(HERE, 'msda_ref.npz'), **out)


def msda_bwd_ref():
    """msda_bwd_ref.npz: gradients of the reference's own ms_deform_attn_core_pytorch (ops/functions/ms_deform_attn_func.py:
    41-61, executed unmodified) by autograd in float64 — what the reference's gradient test (ops/test.py:63-86,
    check_gradient_numerical) holds its CUDA backward against."""
    src = open(os.path.join(REF, 'ops/functions/ms_deform_attn_func.py')).read().split('\n')
    ns = {'torch': torch, 'F': torch.nn.functional}
    exec('\n'.join(src[40:61]), ns)
    core = ns['ms_deform_attn_core_pytorch']
    out = {}
    g = torch.Generator().manual_seed(11)
    cases = {
        'a': dict(N=1, M=2, D=2, Lq=2, shapes=[(6, 4), (3, 2)], P=2, spread=1.0),        # ops/test.py:20-36
        'b': dict(N=2, M=8, D=16, Lq=50, shapes=[(12, 14)], P=16, spread=1.2),           # IGF layout, samples outside [0,1]
        'c': dict(N=1, M=3, D=5, Lq=7, shapes=[(5, 7), (4, 3), (2, 2)], P=3, spread=1.1),   # D not a power of two
    }
    for k, c in cases.items():
        shapes = torch.as_tensor(c['shapes'], dtype=torch.long)
        S = int((shapes[:, 0] * shapes[:, 1]).sum())
        Lv = len(c['shapes'])
        value = torch.randn(c['N'], S, c['M'], c['D'], generator=g)
        loc = torch.rand(c['N'], c['Lq'], c['M'], Lv, c['P'], 2, generator=g) * c['spread'] - (c['spread'] - 1) / 2
        aw = torch.softmax(torch.randn(c['N'], c['Lq'], c['M'], Lv * c['P'], generator=g), -1).view(c['N'], c['Lq'], c['M'], Lv, c['P'])
        go = torch.randn(c['N'], c['Lq'], c['M'] * c['D'], generator=g)
        v64, l64, a64 = (t.double().requires_grad_(True) for t in (value, loc, aw))
        y = core(v64, shapes, l64, a64)
        gv, gl, ga = torch.autograd.grad(y, (v64, l64, a64), go.double())
        out.update({k + '_value': value.numpy(), k + '_shapes': shapes.numpy(), k + '_loc': loc.numpy(), k + '_aw': aw.numpy(),
                    k + '_grad_out': go.numpy(), k + '_out': y.detach().float().numpy(), k + '_grad_value': gv.float().numpy(),
                    k + '_grad_loc': gl.float().numpy(), k + '_grad_aw': ga.float().numpy()})
    np.savez_compressed(os.path.join(HERE, 'msda_bwd_ref.npz'), **out)


SPCONV_CASES = {   # the four rulebook shapes of the SparseEncoder (config:72-82; sparse_encoder.py:76-105)
    'subm3': dict(ksize=(3, 3, 3), stride=(1, 1, 1), padding=(1, 1, 1), subm=True),
    'down_p1': dict(ksize=(3, 3, 3), stride=(2, 2, 2), padding=(1, 1, 1), subm=False),
    'down_p011': dict(ksize=(3, 3, 3), stride=(2, 2, 2), padding=(0, 1, 1), subm=False),
    'conv_out': dict(ksize=(3, 1, 1), stride=(2, 1, 1), padding=(0, 0, 0), subm=False),
}
SPCONV_SHAPE, SPCONV_BATCH = [11, 20, 18], 2


def canonical_rulebook(out_indices, pairs_k_in_out, out_shape):
    """(sorted output coordinates, (k, in_row, rank of the output site) triples sorted) — independent of row order."""
    oi = np.asarray(out_indices, dtype=np.int64)
    key = ((oi[:, 0] * out_shape[0] + oi[:, 1]) * out_shape[1] + oi[:, 2]) * out_shape[2] + oi[:, 3]
    order = np.argsort(key, kind='stable')
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    t = np.asarray(pairs_k_in_out, dtype=np.int64).reshape(-1, 3)
    t = np.stack([t[:, 0], t[:, 1], rank[t[:, 2]]], 1) if len(t) else t
    t = t[np.lexsort((t[:, 2], t[:, 1], t[:, 0]))] if len(t) else t
    return oi[order].astype(np.int32), t.astype(np.int32)


def spconv_ref_cpp():
    """spconv_ref_cpp.npz: rulebooks of the reference's vendored spconv-1.x CPU generator (geometry.h:144-297, compiled by
    oracle/build_ref.py::build_spconv) on a seeded clustered voxel set, in canonical (order-free) form."""
    from oracle import build_ref, ref_spconv
    build_ref.build_spconv()
    rs = np.random.RandomState(5)
    cells = SPCONV_BATCH * SPCONV_SHAPE[0] * SPCONV_SHAPE[1] * SPCONV_SHAPE[2]
    lin = rs.permutation(cells)[:700]
    s0, s1, s2 = SPCONV_SHAPE
    idx = np.stack([lin // (s0 * s1 * s2), (lin // (s1 * s2)) % s0, (lin // s2) % s1, lin % s2], 1).astype(np.int32)
    out = dict(indices=idx)
    for name, c in SPCONV_CASES.items():
        oi, oshape, pairs, num = ref_spconv.indice_pairs(idx, SPCONV_BATCH, SPCONV_SHAPE, c['ksize'], c['stride'], c['padding'],
                                                         (1, 1, 1), c['subm'])
        trip = [(k, pairs[k, 0, j], pairs[k, 1, j]) for k in range(pairs.shape[0]) for j in range(num[k])]
        co, ct = canonical_rulebook(oi, trip, oshape)
        out[name + '_out'], out[name + '_pairs'], out[name + '_shape'] = co, ct, np.asarray(oshape, np.int32)
        print(name, 'sites', len(co), 'pairs', len(ct), 'out shape', oshape)
    np.savez_compressed(os.path.join(HERE, 'spconv_ref_cpp.npz'), **out)


if __name__ == '__main__':
    for name in (sys.argv[1:] or ['voxel_generator_kat', 'voxel_ref_cpp', 'msda_ref', 'msda_bwd_ref', 'spconv_ref_cpp']):
        globals()[name]()
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
