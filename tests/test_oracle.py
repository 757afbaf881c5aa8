"""Self-consistency of the CPU oracle (oracle/me_oracle.py) against stock PyTorch dense operators — the substitute for the
golden vectors the reference does not ship for its MinkowskiEngine boundary.

A sparse op restricted to the active sites must equal the dense op on the densified grid:
  sparse conv            == F.conv3d             (weights permuted from ME's x-fastest [K,Cin,Cout] to [Cout,Cin,kz,ky,kx])
  sparse conv, stride 2  == F.conv3d(stride=2)   (even kernel anchored at the output coordinate)
  sparse conv-transpose  == F.conv_transpose3d   restricted to the encoder's active sites
  avg pooling            == dense sum / dense count
  instance norm          == F.instance_norm per scene (eps 1e-8, biased variance)
  kernel map             == brute-force dictionary lookup
CPU only; sizes chosen so the file runs in seconds.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import me_oracle as O

G = 12  # dense grid edge


def _scene(seed, n=260, batch=2, lo=0, hi=G):
    rng = np.random.default_rng(seed)
    cs = []
    for b in range(batch):
        c = rng.integers(lo, hi, size=(n, 3))
        c = np.unique(c, axis=0)  # lexicographic + unique, as mask3d/datasets/utils.py:106 produces
        cs.append(np.concatenate([np.full((len(c), 1), b), c], 1))
    return np.concatenate(cs).astype(np.int32)


def _densify(coords, feats, grid=G, origin=0):
    B = int(coords[:, 0].max()) + 1
    d = torch.zeros((B, feats.shape[1], grid, grid, grid), dtype=feats.dtype)  # [B, C, z, y, x]
    c = torch.from_numpy(coords.astype(np.int64))
    d[c[:, 0], :, c[:, 3] - origin, c[:, 2] - origin, c[:, 1] - origin] = feats
    return d


def _sample(dense, coords, scale=1, origin=0):
    c = torch.from_numpy(coords.astype(np.int64))
    return dense[c[:, 0], :, (c[:, 3] - origin) // scale, (c[:, 2] - origin) // scale, (c[:, 1] - origin) // scale]


def _dense_weight(W, ks):
    """ME layout [K, Cin, Cout] with k = ix + Kx*(iy + Ky*iz)  ->  conv3d layout [Cout, Cin, kz, ky, kx]."""
    K, cin, cout = W.shape
    return W.reshape(ks, ks, ks, cin, cout).permute(4, 3, 0, 1, 2).contiguous()


@pytest.mark.parametrize("ks", [3, 5])
def test_sparse_conv_equals_dense_conv3d(ks):
    coords = _scene(ks)
    x = torch.randn(len(coords), 5, dtype=torch.float64)
    W = torch.randn(ks ** 3, 5, 7, dtype=torch.float64)
    nbr = O.neighbor_table(coords, coords, O.kernel_offsets((ks,) * 3, (1,) * 3, (1,) * 3))
    got = O.conv_from_table(x, W, nbr, len(coords))
    dense = F.conv3d(_densify(coords, x), _dense_weight(W, ks), padding=ks // 2)
    assert torch.allclose(got, _sample(dense, coords), atol=1e-10)


def test_strided_conv_equals_dense_stride2():
    coords = _scene(11)
    out_coords = O.stride_coords(coords, (2, 2, 2))
    x = torch.randn(len(coords), 4, dtype=torch.float64)
    W = torch.randn(8, 4, 6, dtype=torch.float64)
    nbr = O.neighbor_table(coords, out_coords, O.kernel_offsets((2,) * 3, (1,) * 3, (1,) * 3))
    got = O.conv_from_table(x, W, nbr, len(out_coords))
    dense = F.conv3d(_densify(coords, x), _dense_weight(W, 2), stride=2)  # window anchored at the even coordinate
    assert torch.allclose(got, _sample(dense, out_coords, scale=2), atol=1e-10)
    # every fine row has exactly one (offset, parent) pair
    assert (np.sort(nbr[nbr >= 0]) == np.arange(len(coords))).all()


def test_conv_transpose_equals_dense_conv_transpose3d():
    fine = _scene(12)
    coarse = O.stride_coords(fine, (2, 2, 2))
    x = torch.randn(len(coarse), 4, dtype=torch.float64)
    W = torch.randn(8, 4, 3, dtype=torch.float64)
    nbr_down = O.neighbor_table(fine, coarse, O.kernel_offsets((2,) * 3, (1,) * 3, (1,) * 3))
    got = O.conv_transpose_from_table(x, W, nbr_down, len(fine))
    cd = torch.zeros((2, 4, G // 2, G // 2, G // 2), dtype=torch.float64)
    c = torch.from_numpy(coarse.astype(np.int64))
    cd[c[:, 0], :, c[:, 3] // 2, c[:, 2] // 2, c[:, 1] // 2] = x
    wd = W.reshape(2, 2, 2, 4, 3).permute(3, 4, 0, 1, 2).contiguous()  # conv_transpose3d: [Cin, Cout, kz, ky, kx]
    dense = F.conv_transpose3d(cd, wd, stride=2)
    assert torch.allclose(got, _sample(dense, fine), atol=1e-10)


def test_avg_pool_equals_dense_sum_over_count():
    coords = _scene(13)
    out_coords = O.stride_coords(coords, (2, 2, 2))
    x = torch.randn(len(coords), 3, dtype=torch.float64)
    nbr = O.neighbor_table(coords, out_coords, O.kernel_offsets((2,) * 3, (1,) * 3, (1,) * 3))
    got = O.avg_pool_from_table(x, nbr, len(out_coords), average=True)
    s = F.avg_pool3d(_densify(coords, x), 2, divisor_override=1)
    cnt = F.avg_pool3d(_densify(coords, torch.ones(len(coords), 1, dtype=torch.float64)), 2, divisor_override=1)
    assert torch.allclose(got, _sample(s, out_coords, 2) / _sample(cnt, out_coords, 2), atol=1e-12)
    assert torch.allclose(O.avg_pool_from_table(x, nbr, len(out_coords), average=False), _sample(s, out_coords, 2), atol=1e-12)


def test_instance_norm_equals_torch_instance_norm():
    coords = _scene(14)
    x = torch.randn(len(coords), 6, dtype=torch.float64) * 3 + 1
    rows = [np.nonzero(coords[:, 0] == b)[0] for b in range(2)]
    w, b = torch.randn(1, 6, dtype=torch.float64), torch.randn(1, 6, dtype=torch.float64)
    got = O.instance_norm(x, rows, w, b)
    for r in rows:
        ref = F.instance_norm(x[r].t()[None], eps=1e-8)[0].t() * w + b
        assert torch.allclose(got[r], ref, atol=1e-10)


@pytest.mark.parametrize("ks,stride,lo", [(3, 1, 0), (2, 2, 0), (3, 1, -6), (2, 2, -6)])
def test_kernel_map_equals_bruteforce(ks, stride, lo):
    coords = _scene(20 + ks, n=150, lo=lo, hi=lo + G)
    out_coords = O.stride_coords(coords, (stride,) * 3) if stride > 1 else coords
    offs = O.kernel_offsets((ks,) * 3, (1,) * 3, (1,) * 3)
    nbr = O.neighbor_table(coords, out_coords, offs)
    table = {tuple(c): i for i, c in enumerate(coords.tolist())}
    for k, off in enumerate(offs.tolist()):
        for o, c in enumerate(out_coords.tolist()):
            q = (c[0], c[1] + off[0], c[2] + off[1], c[3] + off[2])
            assert nbr[k, o] == table.get(q, -1)
    # negative coordinates floor towards -inf (ME floors in float)
    if stride > 1:
        assert (out_coords[:, 1:] % stride == 0).all()
        assert {tuple(r) for r in out_coords.tolist()} == {(c[0], *(int(np.floor(v / stride)) * stride for v in c[1:])) for c in coords.tolist()}


def test_insert_and_map_first_occurrence_order():
    c = np.array([[0, 5, 5, 5], [0, 1, 1, 1], [0, 5, 5, 5], [1, 1, 1, 1], [0, 1, 1, 1], [0, 0, 0, 9]], dtype=np.int32)
    uniq, inv = O.insert_and_map(c)
    assert uniq.tolist() == [0, 1, 3, 5] and inv.tolist() == [0, 1, 0, 2, 1, 3]


def test_sparse_quantize_labels_and_floor():
    pts = np.array([[0.1, 0.2, 0.3], [0.4, 0.1, 0.2], [-0.1, 0.0, 0.0], [1.2, 0.0, 0.0]])
    lab = np.array([3, 4, 5, 6])
    c, l, idx, inv = O.sparse_quantize(pts, labels=lab, return_index=True, return_inverse=True, quantization_size=0.5, ignore_label=-100)
    assert c.tolist() == [[0, 0, 0], [-1, 0, 0], [2, 0, 0]]
    assert l.tolist() == [-100, 5, 6] and idx.tolist() == [0, 2, 3] and inv.tolist() == [0, 0, 1, 2]
