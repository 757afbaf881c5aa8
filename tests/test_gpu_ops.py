"""GPU parity: every kernel family of libmk3d.so (called through the C ABI) against the CPU oracle on seeded inputs.

Tolerances (BASELINE.json north_star): maps / indices bit-exact; conv, norm, pooling, mask logits, attention within
1e-3 relative in fp32 mode; bf16 tensor-core mode is compared (a) against the oracle fed the SAME bf16-rounded operands
(tests the kernel: tolerance 2e-3 of the output scale, fp32 accumulation order differs) and (b) against the fp32 oracle
with the stated bf16 tolerance 3e-2.
"""
import numpy as np
import pytest
import torch

from oracle import fps_oracle, me_oracle as O
from room_instance_segmentation_mask3d_b200 import scenes

pytestmark = pytest.mark.gpu


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def R():
    from room_instance_segmentation_mask3d_b200 import engine, lib, me_api, ops

    lib.load()

    class NS:
        pass

    ns = NS()
    ns.engine, ns.lib, ns.me, ns.ops = engine, lib, me_api, ops
    return ns


def make_batch(seed=0, with_dups=False):
    sc = [scenes.tiny_scene(seed), scenes.tiny_scene(seed + 1, (20, 16, 10)), scenes.tiny_scene(seed + 2, (9, 30, 7), 1)]
    coords, feats = scenes.collate(sc)
    if with_dups:
        rng = np.random.default_rng(seed)
        dup = rng.integers(0, len(coords), size=len(coords) // 3)
        coords = np.concatenate([coords, coords[dup]], 0)
        order = np.argsort(coords[:, 0], kind="stable")  # keep batches contiguous, duplicates interleaved at the end of each
        coords = coords[order]
    return coords


# ------------------------------------------------------------------------------------------ coordinate maps
@pytest.mark.parametrize("with_dups", [False, True])
def test_insert_and_map_bit_exact(R, with_dups):
    coords = make_batch(3, with_dups)
    mgr = R.engine.CoordinateManager(D=3)
    key, (uniq, inv) = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    u_ref, inv_ref = O.insert_and_map(coords)
    assert np.array_equal(uniq.cpu().numpy(), u_ref)
    assert np.array_equal(inv.cpu().numpy(), inv_ref)
    assert np.array_equal(mgr.get_coordinates(key).cpu().numpy(), coords[u_ref])


def test_empty_and_single(R):
    mgr = R.engine.CoordinateManager(D=3)
    key, (uniq, inv) = mgr.insert_and_map(torch.zeros((0, 4), dtype=torch.int32).cuda())
    assert mgr.size(key) == 0 and uniq.numel() == 0
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.tensor([[0, -5, 7, 9]], dtype=torch.int32).cuda())
    assert mgr.size(key) == 1


def test_out_of_range_coordinate_raises(R):
    mgr = R.engine.CoordinateManager(D=3)
    with pytest.raises(RuntimeError):
        mgr.insert_and_map(torch.tensor([[0, 1 << 17, 0, 0]], dtype=torch.int32).cuda())


def test_voxelize_collate_matches_numpy_unique(R):
    """GPU dedup + collate == the reference's np.unique(axis=0, return_index=True) per scene + sparse_collate
    (mask3d/datasets/utils.py:96-126): same rows, same lexicographic order, same first-occurrence indices; negative
    coordinates, duplicates (augmentation rounding), float inputs truncated like `.int()`, an empty scene in the middle."""
    rng = np.random.default_rng(3)
    scenes_in, feats_in, labels_in = [], [], []
    for b, n in enumerate([5000, 0, 1237, 20000]):
        c = rng.integers(-40, 40, size=(n, 3)).astype(np.float32) + rng.uniform(-0.4, 0.4, size=(n, 3)).astype(np.float32) * (b % 2)
        if n:
            c = np.concatenate([c, c[rng.integers(0, n, size=n // 4)]], 0)  # exact duplicates
        scenes_in.append(c)
        feats_in.append(rng.standard_normal((len(c), 3)).astype(np.float32))
        labels_in.append(rng.integers(0, 7, size=(len(c), 2)))
    bcoords, feats, labels, uidx, (mgr, key) = R.me.voxelize_collate(scenes_in, feats_in, labels_in)
    want_c, want_f, want_l = [], [], []
    for b, (c, f, l) in enumerate(zip(scenes_in, feats_in, labels_in)):
        ci = torch.from_numpy(c).int().numpy()
        u, idx = np.unique(ci, axis=0, return_index=True)
        assert np.array_equal(uidx[b].cpu().numpy(), idx), f"scene {b}: first-occurrence indices"
        want_c.append(np.concatenate([np.full((len(u), 1), b, dtype=np.int32), u.astype(np.int32)], 1))
        want_f.append(f[idx])
        want_l.append(l[idx])
    assert np.array_equal(bcoords.cpu().numpy(), np.concatenate(want_c, 0))
    assert np.array_equal(feats.cpu().numpy(), np.concatenate(want_f, 0))
    assert np.array_equal(labels.cpu().numpy(), np.concatenate(want_l, 0))
    # the coordinate map built on the way is the SparseTensor's map
    x = R.me.SparseTensor(feats, coordinate_map_key=key, coordinate_manager=mgr)
    assert x.C.shape[0] == feats.shape[0] and np.array_equal(x.C.cpu().numpy(), bcoords.cpu().numpy())


def test_stride_maps_bit_exact(R):
    coords = make_batch(5)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    ref = coords
    for lvl in range(1, 5):
        key = mgr.stride(key, 2)
        ts = (2 ** lvl,) * 3
        ref = O.stride_coords(ref, ts)
        assert key.tensor_stride == ts
        assert np.array_equal(mgr.get_coordinates(key).cpu().numpy(), ref), f"level {lvl}"


@pytest.mark.parametrize("ksize,stride", [(3, 1), (5, 1), (2, 2), (1, 1)])
def test_kernel_map_bit_exact(R, ksize, stride):
    coords = make_batch(7)
    mgr = R.engine.CoordinateManager(D=3)
    k_in, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    k_out = mgr.stride(k_in, stride)
    km = mgr.neighbor_table(k_in, k_out, ksize, 1)
    out_coords = mgr.get_coordinates(k_out).cpu().numpy()
    ref = O.neighbor_table(coords, out_coords, O.kernel_offsets((ksize,) * 3, (1,) * 3, (1,) * 3))
    assert np.array_equal(km.nbr.cpu().numpy(), ref)
    assert km.num_pairs == int((ref >= 0).sum())
    # ME-visible dict form (canonical: sorted by out row)
    d = mgr.kernel_map(k_in, k_out, stride=stride, kernel_size=ksize)
    dref = O.kernel_map_from_table(ref)
    assert sorted(d) == sorted(dref)
    for k in dref:
        assert np.array_equal(d[k].cpu().numpy(), dref[k])
    # per-tile offset masks
    if ksize ** 3 <= 32:
        kmask = km.tile_kmask.cpu().numpy().view(np.uint32)
        for t in range(len(kmask)):
            sl = ref[:, t * 128:(t + 1) * 128]
            want = sum(1 << k for k in range(ref.shape[0]) if (sl[k] >= 0).any())
            assert int(kmask[t]) == want


def test_kernel_map_on_strided_level(R):
    coords = make_batch(9)
    mgr = R.engine.CoordinateManager(D=3)
    k0, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    k1 = mgr.stride(k0, 2)
    k2 = mgr.stride(k1, 2)
    c1, c2 = mgr.get_coordinates(k1).cpu().numpy(), mgr.get_coordinates(k2).cpu().numpy()
    km = mgr.neighbor_table(k1, k1, 3, 1)
    assert np.array_equal(km.nbr.cpu().numpy(), O.neighbor_table(c1, c1, O.kernel_offsets((3,) * 3, (1,) * 3, (2,) * 3)))
    km = mgr.neighbor_table(k1, k2, 2, 1)
    assert np.array_equal(km.nbr.cpu().numpy(), O.neighbor_table(c1, c2, O.kernel_offsets((2,) * 3, (1,) * 3, (2,) * 3)))


def test_hash_query(R):
    coords = make_batch(1)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    cm = mgr.coord_map(key)
    q = np.concatenate([coords[::7], coords[::11] + np.array([0, 1000, 0, 0], dtype=np.int32)])
    rows = torch.empty(len(q), dtype=torch.int32, device="cuda")
    R.lib.call("mk_hash_query", R.lib.p(cm.keys), R.lib.p(cm.vals), cm.cap, R.lib.p(torch.from_numpy(q).cuda()), len(q), R.lib.p(rows), R.lib.stream())
    want = np.concatenate([np.arange(len(coords))[::7], -np.ones(len(coords[::11]), dtype=np.int64)])
    assert np.array_equal(rows.cpu().numpy(), want)


# ------------------------------------------------------------------------------------------ convolution
def _conv_setup(R, cin, cout, ksize, stride=1, seed=0):
    coords = make_batch(seed)
    mgr = R.engine.CoordinateManager(D=3)
    k_in, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    k_out = mgr.stride(k_in, stride)
    km = mgr.neighbor_table(k_in, k_out, ksize, 1)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(len(coords), cin, generator=g)
    W = torch.randn(ksize ** 3, cin, cout, generator=g) / np.sqrt(cin * ksize ** 3)
    bias = torch.randn(1, cout, generator=g) * 0.1
    return mgr, k_in, k_out, km, x, W, bias


@pytest.mark.parametrize("cin,cout,ksize,stride", [(32, 32, 3, 1), (96, 96, 3, 1), (128, 96, 3, 1), (32, 64, 2, 2), (48, 16, 3, 1)])
def test_conv_fp32_parity(R, cin, cout, ksize, stride):
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, cin, cout, ksize, stride)
    ref = O.conv_from_table(x, W, km.nbr.cpu().numpy(), km.n_out) + bias
    out = R.ops.conv_forward(x.cuda(), W.cuda(), km, bias=bias.cuda(), mode="fp32")
    assert rel_err(out, ref) < 1e-5


@pytest.mark.parametrize("cin,cout,ksize,stride", [(32, 32, 3, 1), (64, 64, 3, 1), (96, 96, 3, 1), (128, 96, 3, 1), (192, 128, 3, 1),
                                                  (384, 256, 3, 1), (256, 256, 3, 1), (32, 32, 2, 2), (128, 128, 2, 2), (96, 128, 1, 1),
                                                  (16, 16, 3, 1)])
def test_conv_bf16_tensor_core_parity(R, cin, cout, ksize, stride):
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, cin, cout, ksize, stride, seed=cin + cout)
    xb, Wb = x.bfloat16().float(), W.bfloat16().float()
    ref_same_operands = O.conv_from_table(xb, Wb, km.nbr.cpu().numpy(), km.n_out) + bias
    ref_fp32 = O.conv_from_table(x, W, km.nbr.cpu().numpy(), km.n_out) + bias
    out = R.ops.conv_forward(x.cuda(), W.cuda(), km, bias=bias.cuda(), mode="bf16", out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert rel_err(out, ref_same_operands) < 2e-3, "tcgen05 kernel disagrees with the oracle on identical bf16 operands"
    assert rel_err(out, ref_fp32) < 3e-2
    out_b = R.ops.conv_forward(x.cuda(), W.cuda(), km, bias=bias.cuda(), mode="bf16", out_dtype=torch.bfloat16)
    assert rel_err(out_b.float(), ref_same_operands) < 1e-2


def test_rows_not_sorted_by_batch_index_raise(R):
    """scene row ranges assume batch-sorted rows (sparse_collate's output); anything else must fail loudly, not silently mis-segment"""
    c = torch.tensor([[1, 0, 0, 0], [0, 1, 0, 0], [1, 2, 0, 0]], dtype=torch.int32).cuda()
    with pytest.raises(RuntimeError, match="sorted by batch"):
        R.engine.CoordinateManager(D=3).insert_and_map(c)


def _tf32_round(t):
    """round-to-nearest-even to TF32 (10 explicit mantissa bits): what the tensor core keeps of an fp32 operand at best"""
    i = t.contiguous().view(torch.int32)
    r = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return r.view(torch.float32)


@pytest.mark.parametrize("cin,cout,ksize,stride", [(32, 32, 3, 1), (64, 64, 3, 1), (96, 96, 3, 1), (128, 96, 3, 1), (192, 128, 3, 1),
                                                  (384, 256, 3, 1), (256, 256, 3, 1), (32, 32, 2, 2), (128, 128, 2, 2), (96, 128, 1, 1),
                                                  (8, 16, 3, 1), (40, 16, 3, 1)])
def test_conv_tf32_tensor_core_parity(R, cin, cout, ksize, stride):
    """tcgen05.mma.kind::tf32 on fp32 activations (the reference's own precision, main_instance_segmentation.py:52) against the
    fp32 oracle: <= 1e-3 relative (BASELINE.json north_star tolerance for the fp32-accumulate mode)."""
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, cin, cout, ksize, stride, seed=cin + cout + 3)
    ref_fp32 = O.conv_from_table(x, W, km.nbr.cpu().numpy(), km.n_out) + bias
    out = R.ops.conv_forward(x.cuda(), W.cuda(), km, bias=bias.cuda(), mode="tf32")
    torch.cuda.synchronize()
    assert out.dtype == torch.float32
    assert rel_err(out, ref_fp32) < 1e-3
    # on operands that are exactly representable in TF32 only the fp32 accumulation order differs
    xt, Wt = _tf32_round(x), _tf32_round(W)
    ref_same = O.conv_from_table(xt, Wt, km.nbr.cpu().numpy(), km.n_out) + bias
    out_same = R.ops.conv_forward(xt.cuda(), Wt.cuda(), km, bias=bias.cuda(), mode="tf32")
    assert rel_err(out_same, ref_same) < 2e-5


def test_conv_transpose_tf32_parity(R):
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, 64, 64, 2, 2, seed=11)
    plan = mgr.transpose_plan(k_out, k_in, 2, 1)
    g = torch.Generator().manual_seed(3)
    xc = torch.randn(mgr.size(k_out), 64, generator=g)
    Wt = torch.randn(8, 64, 48, generator=g) / 8.0
    want = R.ops.conv_transpose_forward_fp32(xc.cuda(), Wt.cuda(), plan)
    got = R.ops.conv_forward(xc.cuda(), Wt.cuda(), plan, mode="tf32")
    assert rel_err(got, want.cpu()) < 1e-3


def _wgrad_ref(x, dy, nbr, K):
    """dW[k] = sum_o x[nbr[k, o]]^T dy[o] in fp64 (the oracle's per-offset gather-mm, me_oracle.conv_from_table differentiated)."""
    out = torch.zeros(K, x.shape[1], dy.shape[1], dtype=torch.float64)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o):
            out[k] = x[nbr[k, o]].double().t() @ dy[o].double()
    return out


@pytest.mark.parametrize("cin,cout,ksize,stride", [(32, 32, 3, 1), (96, 96, 3, 1), (128, 96, 3, 1), (288, 48, 3, 1), (64, 256, 3, 1),
                                                  (256, 256, 3, 1), (96, 96, 2, 2), (48, 16, 1, 1), (24, 16, 3, 1)])
def test_conv_wgrad_tensor_core(R, cin, cout, ksize, stride):
    """tcgen05 weight gradient (MN-major operands) against the fp64 per-offset product on the SAME bf16-rounded operands."""
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, cin, cout, ksize, stride, seed=cin + cout + 1)
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(km.n_out, cout, generator=g)
    xb, dyb = x.bfloat16(), dy.bfloat16()
    K = ksize ** 3
    ref = _wgrad_ref(xb.float(), dyb.float(), km.nbr.cpu().numpy(), K)
    got = R.ops.conv_wgrad_bf16(xb.cuda(), dyb.cuda(), km, K, cin, cout)
    torch.cuda.synchronize()
    assert rel_err(got, ref) < 1e-4
    # per offset too: a wrong offset -> weight slot mapping could hide in the global maximum
    for k in range(K):
        if float(ref[k].abs().max()) > 0:
            assert rel_err(got[k], ref[k]) < 1e-3, f"offset {k}"
        else:
            assert float(got[k].abs().max()) == 0.0


def test_conv_wgrad_tensor_core_transposed_plan(R):
    """weight gradient of the transposed conv: same kernel over the per-offset padded transpose plan (one offset per tile)."""
    coords = make_batch(4)
    mgr = R.engine.CoordinateManager(D=3)
    k0, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    k1 = mgr.stride(k0, 2)
    plan = mgr.transpose_plan(k1, k0, 2, 1)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(mgr.size(k1), 64, generator=g).bfloat16()
    dy = torch.randn(len(coords), 96, generator=g).bfloat16()
    parent, koff = plan.parent.cpu().numpy(), plan.koff.cpu().numpy()
    ref = torch.zeros(8, 64, 96, dtype=torch.float64)
    for k in range(8):
        f = np.nonzero(koff == k)[0]
        ref[k] = x.float()[parent[f]].double().t() @ dy.float()[f].double()
    got = R.ops.conv_wgrad_bf16(x.cuda(), dy.cuda(), plan, 8, 64, 96)
    assert rel_err(got, ref) < 1e-4


@pytest.mark.parametrize("cin,cout,ksize", [(96, 96, 3), (288, 64, 3), (64, 128, 1)])
def test_conv_dgrad_tensor_core(R, cin, cout, ksize):
    """dgrad = the forward kernel over the same-map table with mirrored offsets and W^T (column slices when cin > 256)."""
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, cin, cout, ksize, 1, seed=cin + cout + 2)
    g = torch.Generator().manual_seed(6)
    dy = torch.randn(km.n_out, cout, generator=g).bfloat16()
    K = ksize ** 3
    nbr = km.nbr.cpu().numpy()
    Wb = W.bfloat16().float()
    ref = torch.zeros(len(nbr[0]), cin, dtype=torch.float64)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o):
            ref.index_add_(0, torch.from_numpy(nbr[k, o]).long(), dy.float()[o].double() @ Wb[k].double().t())
    kidx = torch.arange(K - 1, -1, -1, dtype=torch.int32, device="cuda")
    got = R.ops.conv_dgrad_bf16(dy.cuda(), W.cuda(), km, kidx=kidx)
    assert rel_err(got, ref) < 2e-3


def test_conv_stem_5x5x5(R):
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, 1, 32, 5, 1)
    ref = O.conv_from_table(x, W, km.nbr.cpu().numpy(), km.n_out)
    out = R.ops.conv_forward(x.cuda(), W.cuda(), km, mode="fp32")
    assert rel_err(out, ref) < 1e-5
    out = R.ops.conv_forward(x.cuda(), W.cuda(), km, mode="bf16", out_dtype=torch.float32)
    assert rel_err(out, ref) < 1e-5  # the stem always runs in fp32 on CUDA cores


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("bf16", 3e-2)])
def test_conv_transpose_parity(R, mode, tol):
    coords = make_batch(11)
    mgr = R.engine.CoordinateManager(D=3)
    k0, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    k1 = mgr.stride(k0, 2)
    g = torch.Generator().manual_seed(1)
    n1 = mgr.size(k1)
    x = torch.randn(n1, 64, generator=g)
    W = torch.randn(8, 64, 32, generator=g) / np.sqrt(64 * 8)
    nbr_down = mgr.neighbor_table(k0, k1, 2, 1).nbr.cpu().numpy()
    ref = O.conv_transpose_from_table(x if mode == "fp32" else x.bfloat16().float(), W if mode == "fp32" else W.bfloat16().float(), nbr_down, len(coords))
    plan = mgr.transpose_plan(k1, k0, 2, 1)
    if mode == "fp32":
        out = R.ops.conv_transpose_forward_fp32(x.cuda(), W.cuda(), plan)
    else:
        out = R.ops.conv_forward(x.cuda(), W.cuda(), plan, mode="bf16", out_dtype=torch.float32)
    assert rel_err(out, ref) < (tol if mode == "fp32" else 2e-3)


# ------------------------------------------------------------------------------------------ norm / pool
@pytest.mark.parametrize("C", [32, 96, 256, 384])
def test_instance_norm_parity(R, C):
    coords = make_batch(13)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    g = torch.Generator().manual_seed(C)
    x = torch.randn(len(coords), C, generator=g) * 3 + 5  # large mean: exercises the shifted-sum statistics
    gamma, beta = torch.randn(1, C, generator=g), torch.randn(1, C, generator=g)
    res = torch.randn(len(coords), C, generator=g)
    rows = [np.nonzero(coords[:, 0] == b)[0] for b in range(int(coords[:, 0].max()) + 1)]
    ref = torch.relu(O.instance_norm(x, rows, gamma, beta) + res)
    off, B = mgr.scene_offsets(key), mgr.number_of_unique_batch_indices()
    out = R.ops.instance_norm(x.cuda(), off, B, gamma.cuda(), beta.cuda(), residual=res.cuda(), relu=True)
    assert rel_err(out, ref) < 1e-4
    out_b = R.ops.instance_norm(x.cuda().bfloat16(), off, B, gamma.cuda(), beta.cuda(), relu=False, out_dtype=torch.float32)
    ref_b = O.instance_norm(x.bfloat16().float(), rows, gamma, beta)
    assert rel_err(out_b, ref_b) < 1e-4


@pytest.mark.parametrize("C", [24, 40, 96, 192])
def test_instance_norm_bf16_with_an_empty_scene_forward_and_backward(R, C):
    """The bf16 -> bf16 apply kernel (a thread keeps one channel group; grid rounded to a multiple of the vectors per row), the 1-D
    statistics grid over non-empty (scene, chunk) pairs and the merge kernels, on a batch whose middle scene has NO rows, for
    channel counts whose vectors per row do not divide 256; forward vs a float64 reference, backward vs autograd through it."""
    g = torch.Generator().manual_seed(100 + C)
    sizes = [700, 0, 1300, 37]
    off = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32).cuda()
    B, n = len(sizes), sum(sizes)
    x = (torch.randn(n, C, generator=g) * 2 + 3).bfloat16()
    res = torch.randn(n, C, generator=g).bfloat16()
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    dy = torch.randn(n, C, generator=g).bfloat16()

    def ref_fn(xd, gd, bd):
        outs = []
        for b in range(B):
            xb = xd[off[b]:off[b + 1]]
            if xb.shape[0] == 0:
                continue
            mu, var = xb.mean(0, keepdim=True), xb.var(0, unbiased=False, keepdim=True)
            outs.append((xb - mu) / torch.sqrt(var + 1e-8) * gd + bd)
        return torch.relu(torch.cat(outs) + res.double())

    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    ref = ref_fn(xd, gd, bd)
    ref.backward(dy.double())
    xg = x.cuda()
    mean, rstd = R.ops.instance_norm_stats(xg, off, B)
    out = R.ops.norm_apply(xg, off, B, mean, rstd, gamma.cuda(), beta.cuda(), residual=res.cuda(), relu=True, out_dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16
    assert rel_err(out.float(), ref.detach().float()) < 1e-2  # bf16 output rounding
    dx, dres, dgamma, dbeta = R.ops.instnorm_relu_bwd(dy.cuda(), out, xg, off, B, mean, rstd, gamma.cuda(), want_dres=True)
    # the kernel masks with ITS bf16 output; rows where the reference output is within rounding of zero may differ: compare sums
    assert rel_err(dx.float(), xd.grad.float()) < 3e-2
    assert rel_err(dgamma, gd.grad.float()) < 2e-2 and rel_err(dbeta, bd.grad.float()) < 2e-2
    assert torch.equal(dres.float() != 0, (out.float() > 0) & (dy.cuda().float() != 0))


def test_avg_pool_parity(R):
    coords = make_batch(17)
    mgr = R.engine.CoordinateManager(D=3)
    k0, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    k1 = mgr.stride(k0, 2)
    km = mgr.neighbor_table(k0, k1, 2, 1)
    x = torch.randn(len(coords), 25)
    for avg in (True, False):
        ref = O.avg_pool_from_table(x, km.nbr.cpu().numpy(), km.n_out, average=avg)
        out = R.ops.pool_forward(x.cuda(), km, average=avg)
        assert rel_err(out, ref) < 1e-6


# ------------------------------------------------------------------------------------------ decoder pieces
def test_mask_logits_and_threshold(R):
    coords = make_batch(19)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    off, B = mgr.scene_offsets(key), mgr.number_of_unique_batch_indices()
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(len(coords), 128, generator=g)
    for Q in (25, 100):
        emb = torch.randn(B, Q, 128, generator=g)
        ref = torch.cat([feat[coords[:, 0] == b] @ emb[b].T for b in range(B)])
        out = R.ops.mask_logits(feat.cuda(), off, B, emb.cuda())
        assert rel_err(out, ref) < 1e-5
        blocked = R.ops.threshold_blocked(out)
        assert torch.equal(blocked.cpu().bool(), out.cpu().sigmoid() < 0.5)


def test_mask_logits_tensor_core_path(R):
    """bf16 features: the product runs on the tcgen05 conv kernel with one weight per scene (scene-select table)."""
    coords = make_batch(29)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    off, B = mgr.scene_offsets(key), mgr.number_of_unique_batch_indices()
    g = torch.Generator().manual_seed(4)
    feat = torch.randn(len(coords), 128, generator=g)
    sel = mgr.scene_select_table(key)
    for Q in (25, 100):
        emb = torch.randn(B, Q, 128, generator=g)
        ref = torch.cat([feat.bfloat16().float()[coords[:, 0] == b] @ emb[b].bfloat16().float().T for b in range(B)])
        out = R.ops.mask_logits(feat.cuda().bfloat16(), off, B, emb.cuda(), select=sel)
        assert out.shape == (len(coords), Q) and out.dtype == torch.float32
        assert rel_err(out, ref) < 2e-3


@pytest.mark.parametrize("kind", ["scene", "sparse_random", "strided"])
def test_composed_5x5x5_kernel_map_is_bit_identical_to_probing(R, kind, monkeypatch):
    """mk_kernel_map_compose5 (5^3 table from the 3^3 table of the same map, hash probes only where no intermediate voxel exists)
    == mk_kernel_map_same == the oracle; on surface scenes, on sparse random voxels (most intermediates missing: the fallback
    probes carry the result) and on a stride-2 map (offset step 2)."""
    if kind == "sparse_random":
        rng = np.random.default_rng(5)
        xyz = np.unique(rng.integers(0, 14, size=(900, 3)), axis=0)
        coords = np.concatenate([np.concatenate([np.full((len(xyz), 1), b), xyz + 3 * b], 1) for b in range(2)]).astype(np.int32)
    else:
        coords = make_batch(41)
    outs = []
    for compose in ("1", "0"):
        monkeypatch.setenv("MK3D_KMAP_COMPOSE", compose)
        mgr = R.engine.CoordinateManager(D=3)
        key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
        if kind == "strided":
            key = mgr.stride(key, 2)
        km = mgr.neighbor_table(key, key, 5, 1)
        outs.append((km.nbr.cpu().numpy(), mgr.get_coordinates(key).cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    c = outs[0][1]
    ts = 2 if kind == "strided" else 1
    ref = O.neighbor_table(c, c, O.kernel_offsets((5, 5, 5), (1, 1, 1), (ts, ts, ts)))
    assert np.array_equal(outs[0][0], ref)
    assert (outs[0][0] >= 0).sum() > len(c)  # not a trivial map


def test_stacked_mask_logits_are_bit_identical_to_separate_products(R):
    """mask_logits_stacked: several decoder outputs per pass over the features, every output bit-identical to its own product
    (bf16 features on the tcgen05 kernel; 13 outputs in groups of 6 as the inference forward uses them, and a tf32 group)."""
    coords = make_batch(37)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    off, B = mgr.scene_offsets(key), mgr.number_of_unique_batch_indices()
    sel = mgr.scene_select_table(key)
    g = torch.Generator().manual_seed(8)
    feat = torch.randn(len(coords), 128, generator=g).cuda()
    embs = [torch.randn(B, 25, 128, generator=g).cuda() for _ in range(13)]
    fb = feat.bfloat16()
    stacked = R.ops.mask_logits_stacked(fb, off, B, embs, sel)
    assert len(stacked) == 13
    for e, o in zip(embs, stacked):
        assert o.shape == (len(coords), 25)
        assert torch.equal(o, R.ops.mask_logits(fb, off, B, e, select=sel))
    R.ops.set_compute_mode("tf32")
    try:
        st32 = R.ops.mask_logits_stacked(feat, off, B, embs[:6], sel)
        for e, o in zip(embs[:6], st32):
            assert torch.equal(o, R.ops.mask_logits(feat, off, B, e, select=sel))
    finally:
        R.ops.set_compute_mode("fp32")


def test_sorted_tile_plan_is_a_permutation(R):
    coords = make_batch(31)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    km = mgr.neighbor_table(key, key, 3, 1)
    sp = km.sorted_plan()
    order = sp.out_rows.cpu().numpy()
    assert np.array_equal(np.sort(order), np.arange(km.n_out))
    assert np.array_equal(sp.nbr.cpu().numpy(), km.nbr.cpu().numpy()[:, order])
    masks = ((km.nbr.cpu().numpy() >= 0).astype(np.int64) * (1 << np.arange(27))[:, None]).sum(0)
    assert np.array_equal(masks[order], np.sort(masks)), "slots must be grouped by neighbour mask"
    # fewer active (tile, offset) pairs than the lexicographic plan
    pop = lambda t: sum(bin(int(v) & 0xFFFFFFFF).count("1") for v in t.cpu().numpy())
    assert pop(sp.tile_kmask) <= pop(km.tile_kmask)


def test_fourier_pe_parity(R):
    from oracle.mask3d_oracle import fourier_pos_enc

    coords = make_batch(23)
    mgr = R.engine.CoordinateManager(D=3)
    key, _ = mgr.insert_and_map(torch.from_numpy(coords).cuda())
    off, B = mgr.scene_offsets(key), mgr.number_of_unique_batch_indices()
    xyz = torch.from_numpy(coords[:, 1:]).float()
    gB = torch.randn(3, 64, generator=torch.Generator().manual_seed(5))
    cmin, cmax = R.ops.scene_minmax(xyz.cuda(), off, B)
    pe = R.ops.fourier_pe(xyz.cuda(), off, B, cmin, cmax, gB.cuda())
    for b in range(B):
        sel = coords[:, 0] == b
        c = xyz[sel]
        ref = fourier_pos_enc(c[None], gB, c.min(0)[0][None], c.max(0)[0][None])[0].T
        assert torch.equal(cmin[b].cpu(), c.min(0)[0]) and torch.equal(cmax[b].cpu(), c.max(0)[0])
        assert (pe.cpu()[sel] - ref).abs().max() < 2e-4  # sin/cos of arguments up to ~50 rad: a few ulp of the argument


@pytest.mark.parametrize("Q", [25, 100])
def test_masked_cross_attention_vs_stock_mha(R, Q):
    """Reference formulation: pad to max K, dense bool mask, all-blocked rule, stock MHA (mask3d.py:400-478)."""
    torch.manual_seed(Q)
    D, H, B = 128, 8, 3
    sizes = [700, 37, 1300]
    off = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    n = sum(sizes)
    mha = torch.nn.MultiheadAttention(D, H)
    query = torch.randn(Q, B, D)
    key_in = torch.randn(n, D)
    val_in = torch.randn(n, D)
    blocked = torch.rand(n, Q) < 0.6
    blocked[off[1]:off[2], 3] = True  # query 3 blocks every key of scene 1 -> must attend to all of them
    blocked[:, 5] = True              # query 5 blocks everything everywhere
    Kmax = max(sizes)
    kpad, vpad, mpad = torch.zeros(Kmax, B, D), torch.zeros(Kmax, B, D), torch.ones(B, Kmax, Q, dtype=torch.bool)
    for b in range(B):
        s, e = int(off[b]), int(off[b + 1])
        idx = torch.zeros(Kmax, dtype=torch.long)
        idx[: e - s] = torch.arange(e - s)
        kpad[:, b], vpad[:, b] = key_in[s:e][idx], val_in[s:e][idx]
        m = blocked[s:e][idx].clone()
        mpad[b] = m
    mpad.permute(0, 2, 1)[mpad.sum(1) == Kmax] = False
    pad = torch.stack([torch.arange(Kmax) >= sz for sz in sizes])
    mpad = torch.logical_or(mpad, pad[..., None])
    with torch.no_grad():
        ref, _ = mha(query, kpad, vpad, attn_mask=mpad.repeat_interleave(H, dim=0).permute(0, 2, 1))
        Wq, Wk, Wv = mha.in_proj_weight.chunk(3)
        bq, bk, bv = mha.in_proj_bias.chunk(3)
        qp = (query.permute(1, 0, 2) @ Wq.T + bq).cuda()
        kp = (key_in @ Wk.T + bk).cuda()
        vp = (val_in @ Wv.T + bv).cuda()
        ctx = R.ops.masked_cross_attention(qp, kp, vp, blocked.to(torch.uint8).cuda(), off.cuda(), H)
        out = (ctx.cpu() @ mha.out_proj.weight.T + mha.out_proj.bias).permute(1, 0, 2)
    # NOTE the reference pads with copies of row 0 but masks them, and its all-blocked test counts pads as blocked
    # (sum == Kmax only if every real key is blocked too), so semantics coincide with the varlen kernel.
    assert rel_err(out, ref) < 1e-4


@pytest.mark.parametrize("Q,sizes", [(25, [700, 37, 1300]), (25, [64, 128, 1, 65]), (32, [5000, 333]), (7, [200])])
def test_masked_cross_attention_tensor_core_kernel(R, Q, sizes):
    """bf16 keys / values, D = 128, H = 8, Q <= 32: the flash-attention style mma.sync kernel (k_xattn_partial_mma) against
    (a) the CUDA-core kernel on the same inputs and (b) a float64 dense evaluation of the reference's semantics (mask3d.py:400-478:
    blocked keys get -inf, a query that blocks every key of its scene attends to all of them, softmax over the scene's keys)."""
    torch.manual_seed(Q + len(sizes))
    D, H, B = 128, 8, len(sizes)
    off = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    n = sum(sizes)
    qp = torch.randn(B, Q, D)
    kp = torch.randn(n, D).bfloat16()
    vp = torch.randn(n, D).bfloat16()
    blocked = torch.rand(n, Q) < 0.6
    blocked[off[0]:off[1], min(3, Q - 1)] = True   # a query that blocks every key of scene 0 -> attends to all of them
    blocked[:, min(5, Q - 1)] = True               # a query that blocks everything everywhere
    args = (qp.cuda(), kp.cuda(), vp.cuda(), blocked.to(torch.uint8).cuda(), off.cuda(), H)
    got = R.ops.masked_cross_attention(*args).cpu()
    R.lib.call("mk_set_xattn_simt", 1)
    try:
        simt = R.ops.masked_cross_attention(*args).cpu()
    finally:
        R.lib.call("mk_set_xattn_simt", 0)
    hd = D // H
    want = torch.zeros(B, Q, D, dtype=torch.float64)
    for b in range(B):
        s, e = int(off[b]), int(off[b + 1])
        k, v, m = kp[s:e].double(), vp[s:e].double(), blocked[s:e].clone()
        m[:, m.all(0)] = False
        for h in range(H):
            sc = (qp[b, :, h * hd:(h + 1) * hd].double() @ k[:, h * hd:(h + 1) * hd].T) * hd ** -0.5
            sc = sc.masked_fill(m.T, float("-inf"))
            want[b, :, h * hd:(h + 1) * hd] = torch.softmax(sc, dim=1) @ v[:, h * hd:(h + 1) * hd]
    assert rel_err(simt, want) < 1e-4
    # the tensor-core kernel rounds the scaled queries and the probabilities to bf16 (2^-9 relative each)
    assert rel_err(got, want) < 1e-2
    assert rel_err(got, simt) < 1e-2


@pytest.mark.parametrize("Q,kv_dtype", [(25, torch.float32), (100, torch.float32), (25, torch.bfloat16)])
def test_masked_cross_attention_backward_vs_stock_mha(R, Q, kv_dtype):
    """Gradients through the attention core against autograd through the reference's formulation (pad to max K, dense bool mask,
    all-blocked rule, stock nn.MultiheadAttention, mask3d.py:400-478, 686-705): d/d(query), d/d(key input), d/d(value input) and
    the in-projection weights.  fp32 keys: <= 1e-3 relative; bf16 keys/values (the fused path's dtype): <= 2e-2."""
    torch.manual_seed(100 + Q)
    D, H, B = 128, 8, 3
    sizes = [700, 37, 1300]
    off = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    n = sum(sizes)
    mha = torch.nn.MultiheadAttention(D, H)
    query = torch.randn(Q, B, D, requires_grad=True)
    key_in = torch.randn(n, D, requires_grad=True)
    val_in = torch.randn(n, D, requires_grad=True)
    blocked = torch.rand(n, Q) < 0.6
    blocked[off[1]:off[2], 3] = True
    blocked[:, 5] = True
    Kmax = max(sizes)
    mpad = torch.ones(B, Kmax, Q, dtype=torch.bool)
    kcols, vcols = [], []
    for b in range(B):
        s, e = int(off[b]), int(off[b + 1])
        idx = torch.zeros(Kmax, dtype=torch.long)
        idx[: e - s] = torch.arange(e - s)
        kcols.append(key_in[s:e][idx])
        vcols.append(val_in[s:e][idx])
        mpad[b] = blocked[s:e][idx]
    kpad, vpad = torch.stack(kcols, 1), torch.stack(vcols, 1)
    mpad.permute(0, 2, 1)[mpad.sum(1) == Kmax] = False
    pad = torch.stack([torch.arange(Kmax) >= sz for sz in sizes])
    mpad = torch.logical_or(mpad, pad[..., None])
    ref, _ = mha(query, kpad, vpad, attn_mask=mpad.repeat_interleave(H, dim=0).permute(0, 2, 1))
    gout = torch.randn_like(ref)
    (ref * gout).sum().backward()
    want = [query.grad.clone(), key_in.grad.clone(), val_in.grad.clone(), mha.in_proj_weight.grad.clone()]
    for t in (query, key_in, val_in, mha.in_proj_weight, mha.in_proj_bias, mha.out_proj.weight, mha.out_proj.bias):
        t.grad = None

    m2 = torch.nn.MultiheadAttention(D, H).cuda()
    m2.load_state_dict(mha.state_dict())
    q2, k2, v2 = (t.detach().cuda().requires_grad_(True) for t in (query, key_in, val_in))
    Wq, Wk, Wv = m2.in_proj_weight.chunk(3)
    bq, bk, bv = m2.in_proj_bias.chunk(3)
    qp = q2.permute(1, 0, 2) @ Wq.T + bq
    kp = (k2 @ Wk.T + bk).to(kv_dtype)
    vp = (v2 @ Wv.T + bv).to(kv_dtype)
    ctx = R.ops.masked_cross_attention(qp, kp, vp, blocked.to(torch.uint8).cuda(), off.cuda(), H, max_scene_rows=max(sizes))
    out = (ctx @ m2.out_proj.weight.T + m2.out_proj.bias).permute(1, 0, 2)
    tol_f = 1e-4 if kv_dtype == torch.float32 else 2e-2
    assert rel_err(out.detach().cpu(), ref.detach()) < tol_f
    (out * gout.cuda()).sum().backward()
    got = [q2.grad.cpu(), k2.grad.cpu(), v2.grad.cpu(), m2.in_proj_weight.grad.cpu()]
    tol = 1e-3 if kv_dtype == torch.float32 else 2e-2
    for name, g, w in zip(("query", "key", "value", "in_proj_weight"), got, want):
        assert rel_err(g, w) < tol, (name, rel_err(g, w))


def test_sparse_quantize_matches_oracle(R):
    """ME.utils.sparse_quantize on the GPU hash vs the oracle: floats with a quantisation size, negative
    coordinates, duplicate voxels with agreeing / conflicting labels, numpy and torch inputs, every return combination.
    Voxel rows, first-occurrence indices and inverse maps bit-exact."""
    rng = np.random.default_rng(4)
    pts = rng.uniform(-3.0, 3.0, size=(4000, 3))
    pts[::7] = pts[1::7][: len(pts[::7])]  # exact duplicates
    lab = rng.integers(0, 3, size=len(pts))
    feats = rng.normal(size=(len(pts), 2)).astype(np.float32)
    want = O.sparse_quantize(pts, features=feats, labels=lab, return_index=True, return_inverse=True, quantization_size=0.25, ignore_label=-100)
    got = R.me.sparse_quantize(pts, features=feats, labels=lab, return_index=True, return_inverse=True, quantization_size=0.25, ignore_label=-100)
    assert len(got) == len(want) == 5
    for g, w in zip(got, want):
        assert np.array_equal(np.asarray(g), np.asarray(w))
    assert (np.asarray(got[2]) == -100).any() and (np.asarray(got[2]) != -100).any()
    # torch input, integer coordinates, maps only
    ci = torch.from_numpy(rng.integers(-20, 20, size=(3000, 3)).astype(np.int32))
    m_w, i_w = O.sparse_quantize(ci, return_maps_only=True, return_inverse=True)
    m_g, i_g = R.me.sparse_quantize(ci.cuda(), return_maps_only=True, return_inverse=True)
    assert np.array_equal(m_g.cpu().numpy(), m_w.numpy()) and np.array_equal(i_g.cpu().numpy(), i_w.numpy())
    assert np.array_equal(R.me.sparse_quantize(ci.cuda()).cpu().numpy(), O.sparse_quantize(ci).numpy())


def test_batch_norm_matches_torch_batchnorm1d(R):
    """MinkowskiBatchNorm on the instance-norm kernels vs nn.BatchNorm1d: training output, running statistics, gradients,
    eval output (fp32, <= 1e-4 relative)."""
    ME = R.me
    torch.manual_seed(3)
    n, C = 5000, 48
    coords = torch.cat([torch.zeros(n, 1, dtype=torch.int32), torch.arange(n, dtype=torch.int32)[:, None].repeat(1, 3)], 1).cuda()
    f = (torch.randn(n, C) * 3 + 1).cuda()
    ref = torch.nn.BatchNorm1d(C, eps=1e-5, momentum=0.02).cuda()
    mod = ME.MinkowskiBatchNorm(C, momentum=0.02).cuda()
    with torch.no_grad():
        ref.weight.uniform_(0.5, 1.5); ref.bias.uniform_(-1, 1)
    mod.bn.load_state_dict(ref.state_dict())
    assert sorted(mod.state_dict()) == sorted("bn." + k for k in ref.state_dict())
    for _ in range(2):
        fa, fb = f.clone().requires_grad_(True), f.clone().requires_grad_(True)
        ya = ref(fa)
        yb = mod(ME.SparseTensor(features=fb, coordinates=coords)).F
        assert rel_err(yb, ya) < 1e-4
        g = torch.randn_like(ya)
        ref.zero_grad(); mod.zero_grad()
        (ya * g).sum().backward(); (yb * g).sum().backward()
        assert rel_err(fb.grad, fa.grad) < 1e-4
        assert rel_err(mod.bn.weight.grad, ref.weight.grad) < 1e-4 and rel_err(mod.bn.bias.grad, ref.bias.grad) < 1e-4
        assert rel_err(mod.bn.running_mean, ref.running_mean) < 1e-5 and rel_err(mod.bn.running_var, ref.running_var) < 1e-5
    ref.eval(); mod.eval()
    with torch.no_grad():
        assert rel_err(mod(ME.SparseTensor(features=f, coordinates=coords)).F, ref(f)) < 1e-4
    assert int(mod.bn.num_batches_tracked) == int(ref.num_batches_tracked) == 2


def test_fps_matches_reference_tie_breaking(R):
    sc = [scenes.tiny_scene(0), scenes.tiny_scene(1, (20, 16, 10)), scenes.box_room((1.0, 1.2, 0.6), 0.02), scenes.tiny_scene(5, (7, 5, 4), 0)]
    coords, _ = scenes.collate(sc)
    off = torch.tensor([0] + list(np.cumsum([len(s) for s in sc])), dtype=torch.int32)
    xyz = torch.from_numpy(coords[:, 1:]).float()
    for m in (25, 100):
        idx = R.ops.furthest_point_sample(xyz.cuda(), off.cuda(), len(sc), m).cpu().numpy()
        for b, s in enumerate(sc):
            if len(s) <= m:
                continue
            ref = fps_oracle.furthest_point_sample(s.astype(np.float32), m)
            assert np.array_equal(idx[b], ref), f"scene {b}, m {m}"


def test_mask_compact_bits_and_scores(R):
    """mk_mask_compact: bit q of word n == logits[n, q] > 0; score == sum sigmoid(x)[x>0] / (sum [x>0] + 1e-6) per scene and query
    (trainer/trainer.py:389-392); ragged scenes incl. an empty one, two runs bit-identical."""
    torch.manual_seed(3)
    sizes = [700, 0, 1300, 5]
    Q = 25
    off = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    x = torch.randn(sum(sizes), Q) * 2
    bits, score = R.ops.mask_compact(x.cuda(), off.cuda(), len(sizes))
    bits2, score2 = R.ops.mask_compact(x.cuda(), off.cuda(), len(sizes))
    assert torch.equal(bits, bits2) and torch.equal(score, score2)
    want_bits = ((x > 0).long() << torch.arange(Q)).sum(1).to(torch.int32)
    assert torch.equal(bits.cpu(), want_bits)
    for b, n in enumerate(sizes):
        xs = x[int(off[b]):int(off[b + 1])].double()
        m = xs > 0
        want = (torch.sigmoid(xs) * m).sum(0) / (m.sum(0) + 1e-6)
        assert rel_err(score[b].cpu().double(), want) < 1e-5 or float(want.abs().max()) == 0


@pytest.mark.parametrize("n,distinct", [(1, 1), (31, 3), (1024, 1), (1025, 7), (100_003, 300), (889_057, 1000)])
def test_group_rows_equals_stable_sort(R, n, distinct):
    """mk_group_rows (hash set of distinct keys + counting sort) == the permutation of a stable sort by key."""
    g = torch.Generator().manual_seed(n)
    vals = torch.randint(0, 1 << 27, (distinct,), generator=g, dtype=torch.int64).to(torch.int32)
    # runs of equal keys (as offset masks come) mixed with scattered ones
    idx = torch.randint(0, distinct, (n,), generator=g)
    idx[: n // 2] = torch.sort(idx[: n // 2]).values
    keys = vals[idx].cuda()
    got = R.engine.group_rows(keys)
    want = torch.sort(keys, stable=True).indices.to(torch.int32)
    assert torch.equal(got, want)


def test_group_rows_too_many_distinct_keys_sets_status(R):
    n = 5000
    keys = torch.arange(n, dtype=torch.int32).cuda()  # 5000 distinct keys > 1024
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    wsb = R.lib.query("mk_group_rows_workspace_bytes", n)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    R.lib.call("mk_group_rows", R.lib.p(keys), n, R.lib.p(order), R.lib.p(status), R.lib.p(ws), wsb, R.lib.stream())
    assert int(status.item()) & 1
    assert sorted(order.cpu().tolist()) == list(range(n))  # still a permutation


def test_conv_wgrad_deterministic_mode_is_bit_identical_and_correct(R):
    """mk_set_deterministic(1): the weight gradient is accumulated per work item and merged in a fixed order -- two runs give the
    same bits (the default path adds with red.global in varying order) and the values agree with the fp64 reference."""
    mgr, k_in, k_out, km, x, W, bias = _conv_setup(R, 96, 96, 3, 1, seed=77)
    g = torch.Generator().manual_seed(6)
    dy = torch.randn(km.n_out, 96, generator=g)
    xb, dyb = x.bfloat16(), dy.bfloat16()
    ref = _wgrad_ref(xb.float(), dyb.float(), km.nbr.cpu().numpy(), 27)
    R.ops.set_deterministic(True)
    try:
        a = R.ops.conv_wgrad_bf16(xb.cuda(), dyb.cuda(), km, 27, 96, 96)
        b = R.ops.conv_wgrad_bf16(xb.cuda(), dyb.cuda(), km, 27, 96, 96)
        torch.cuda.synchronize()
    finally:
        R.ops.set_deterministic(False)
    assert torch.equal(a, b)
    assert rel_err(a, ref) < 1e-4
