"""GPU parity of the BACKWARD of the module path (me_api autograd Functions -> fp32 mode: mk_conv_dgrad/wgrad_f32; bf16 mode:
tcgen05 dgrad (mk_conv_fwd_bf16 over the transposed table) + mk_conv_wgrad_bf16; mk_instnorm_bwd, mk_pool_bwd) against torch autograd through the CPU oracle's modules on the same seeded inputs.

The oracle's modules are built from index_select / mm / index_add, so torch differentiates them exactly; the reference's
own backward is ME's ConvolutionBackward + its custom instance-norm backward.
Tolerance: 1e-3 relative (fp32 mode, BASELINE.json north_star); bf16 mode against its stated tolerance."""
import numpy as np
import pytest
import torch

from oracle import me_oracle as O
from room_instance_segmentation_mask3d_b200 import scenes

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ME():
    from room_instance_segmentation_mask3d_b200 import lib, me_api

    lib.load()
    return me_api


def _batch(seed=0):
    sc = [scenes.tiny_scene(seed), scenes.tiny_scene(seed + 1, (20, 16, 10)), scenes.tiny_scene(seed + 2, (9, 30, 7), 1)]
    return scenes.collate(sc)[0]


def _pair(ME, make, coords, cin, seed, pre_stride=1):
    """Build the same module on both sides with identical parameters; returns (gpu_module, oracle_module, x_gpu, x_cpu)."""
    torch.manual_seed(seed)
    m_ref = make(O)
    m_gpu = make(ME)
    m_gpu.load_state_dict(m_ref.state_dict())
    m_gpu = m_gpu.cuda()
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(len(coords), cin, generator=g)
    xr = O.SparseTensor(features=f.clone().requires_grad_(True), coordinates=torch.from_numpy(coords))
    xg = ME.SparseTensor(features=f.clone().cuda().requires_grad_(True), coordinates=torch.from_numpy(coords).cuda())
    return m_gpu, m_ref, xg, xr


def _check(ME, m_gpu, m_ref, xg, xr, mode, tol):
    ME.set_compute_mode(mode)
    try:
        yg, yr = m_gpu(xg), m_ref(xr)
        assert rel(yg.F, yr.F) < tol
        g = torch.Generator().manual_seed(7)
        dy = torch.randn(yr.F.shape, generator=g)
        yr.F.backward(dy)
        yg.F.backward(dy.cuda())
        assert rel(xg.F.grad, xr.F.grad) < tol, "input gradient"
        for (n, pg), (_, pr) in zip(m_gpu.named_parameters(), m_ref.named_parameters()):
            if pr.grad is None:
                assert pg.grad is None or float(pg.grad.abs().max()) == 0.0, n
                continue
            assert rel(pg.grad, pr.grad) < tol, f"gradient of {n}"
    finally:
        ME.set_compute_mode("bf16")


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
@pytest.mark.parametrize("cin,cout,ks,stride", [(16, 32, 3, 1), (32, 32, 2, 2), (48, 16, 1, 1), (96, 96, 3, 1), (288, 48, 3, 1),
                                                (64, 256, 3, 1), (1, 32, 5, 1), (1, 24, 3, 1)])
def test_conv_backward(ME, cin, cout, ks, stride, mode, tol):
    coords = _batch(3)
    mk = lambda M: M.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, dilation=1, bias=True, dimension=3)
    _check(ME, *_pair(ME, mk, coords, cin, seed=cin + ks), mode, tol)


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 3e-2)])
def test_conv_transpose_backward(ME, mode, tol):
    """conv (stride 2) followed by its transposed conv back onto the encoder map, as the U-Net does (res16unet.py:248-276)."""
    import torch.nn as nn

    coords = _batch(5)

    def mk(M):
        class Net(nn.Module):
            def __init__(self):
                super().__init__()
                self.down = M.MinkowskiConvolution(16, 32, kernel_size=2, stride=2, dilation=1, bias=False, dimension=3)
                self.up = M.MinkowskiConvolutionTranspose(32, 16, kernel_size=2, stride=2, dilation=1, bias=False, dimension=3)

            def forward(self, x):
                return self.up(self.down(x))

        return Net()

    _check(ME, *_pair(ME, mk, coords, 16, seed=9), mode, tol)


def test_instance_norm_relu_pool_backward(ME):
    import torch.nn as nn

    coords = _batch(8)

    def mk(M):
        class Net(nn.Module):
            def __init__(self):
                super().__init__()
                self.conv = M.MinkowskiConvolution(8, 16, kernel_size=3, stride=1, dilation=1, bias=False, dimension=3)
                self.norm = M.MinkowskiInstanceNorm(16)
                self.relu = M.MinkowskiReLU(inplace=True)
                self.pool = M.MinkowskiAvgPooling(kernel_size=2, stride=2, dimension=3)

            def forward(self, x):
                return self.pool(self.relu(self.norm(self.conv(x))))

        return Net()

    m_gpu, m_ref, xg, xr = _pair(ME, mk, coords, 8, seed=21)
    with torch.no_grad():  # non-trivial affine parameters
        for m in (m_gpu, m_ref):
            m.norm.weight.copy_(torch.linspace(0.5, 1.5, 16).view(1, 16))
            m.norm.bias.copy_(torch.linspace(-0.2, 0.2, 16).view(1, 16))
    _check(ME, m_gpu, m_ref, xg, xr, "fp32", 1e-3)


def test_basic_block_backward_fp32(ME):
    """One residual block of the backbone (resnet_block.py:52-68) incl. the 1x1 downsample branch, fwd + bwd."""
    import torch.nn as nn

    coords = _batch(13)

    def mk(M):
        class Block(nn.Module):
            def __init__(self):
                super().__init__()
                c = lambda i, o, k: M.MinkowskiConvolution(i, o, kernel_size=k, stride=1, dilation=1, bias=False, dimension=3)
                self.conv1, self.norm1 = c(16, 32, 3), M.MinkowskiInstanceNorm(32)
                self.conv2, self.norm2 = c(32, 32, 3), M.MinkowskiInstanceNorm(32)
                self.ds, self.dsn = c(16, 32, 1), M.MinkowskiInstanceNorm(32)
                self.relu = M.MinkowskiReLU(inplace=True)

            def forward(self, x):
                out = self.relu(self.norm1(self.conv1(x)))
                out = self.norm2(self.conv2(out))
                out += self.dsn(self.ds(x))
                return self.relu(out)

        return Block()

    _check(ME, *_pair(ME, mk, coords, 16, seed=31), "fp32", 1e-3)


# ------------------------------------------------------------------------------------------ fused training unit
def _cos(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


@pytest.mark.parametrize("relu,with_res,ks,stride", [(True, False, 3, 1), (True, True, 3, 1), (False, False, 1, 1), (True, False, 2, 2)])
def test_fused_conv_norm_act_unit(ME, relu, with_res, ks, stride):
    """_ConvNormActFn (one autograd node, bf16) against the fp32 module composition conv -> IN -> (+res) -> (ReLU) on the same
    parameters: outputs and every gradient (input, residual, kernel, gamma, beta)."""
    import torch.nn.functional as F
    from room_instance_segmentation_mask3d_b200 import mask3d_model as MM, ops

    coords = torch.from_numpy(_batch(17)).cuda()
    cin, cout = 32, 48
    torch.manual_seed(3)
    conv = ME.MinkowskiConvolution(cin, cout, kernel_size=ks, stride=stride, dilation=1, bias=False, dimension=3).cuda()
    norm = ME.MinkowskiInstanceNorm(cout).cuda()
    with torch.no_grad():
        norm.weight.copy_(torch.linspace(0.5, 1.5, cout).view(1, cout))
        norm.bias.copy_(torch.linspace(-0.3, 0.3, cout).view(1, cout))
    x0 = torch.randn(coords.shape[0], cin, device="cuda")
    # ---- reference: fp32 modules
    ME.set_compute_mode("fp32")
    xs = ME.SparseTensor(x0.clone().requires_grad_(True), coordinates=coords)
    y = norm(conv(xs))
    n_out = y.F.shape[0]
    res0 = torch.randn(n_out, cout, device="cuda") if with_res else None
    res_ref = res0.clone().requires_grad_(True) if with_res else None
    yf = y.F + res_ref if with_res else y.F
    yf = F.relu(yf) if relu else yf
    dy = torch.randn_like(yf)
    yf.backward(dy)
    ref = dict(y=yf.detach(), dx=xs.F.grad, dW=conv.kernel.grad.clone(), dg=norm.weight.grad.clone(), db=norm.bias.grad.clone(),
               dres=res_ref.grad if with_res else None)
    conv.kernel.grad = norm.weight.grad = norm.bias.grad = None
    # ---- fused unit, bf16
    ME.set_compute_mode("bf16")
    mgr, k_in = xs.coordinate_manager, xs.coordinate_map_key
    k_out = mgr.stride(k_in, stride)
    km = mgr.neighbor_table(k_in, k_out, ks, 1)
    kind, aux = (0, None) if stride == 1 else (1, mgr.transpose_plan(k_out, k_in, ks, 1))
    B = mgr.number_of_unique_batch_indices()
    xb = x0.bfloat16().requires_grad_(True)
    rb = res0.bfloat16().requires_grad_(True) if with_res else None
    out = MM._ConvNormActFn.apply(xb, conv.kernel, norm.weight, norm.bias, rb, km, kind, aux, mgr.scene_offsets(k_out), B, relu)
    out.backward(dy.bfloat16())
    torch.cuda.synchronize()
    assert rel(out.float(), ref["y"]) < 3e-2
    for name, got, want in [("dx", xb.grad, ref["dx"]), ("dW", conv.kernel.grad, ref["dW"]), ("dgamma", norm.weight.grad, ref["dg"]),
                            ("dbeta", norm.bias.grad, ref["db"])] + ([("dres", rb.grad, ref["dres"])] if with_res else []):
        assert _cos(got.float(), want) > 0.999, name
        if not relu:
            assert rel(got.float(), want) < 3e-2, name  # max-norm bound
        else:
            # with a ReLU a handful of pre-activations within bf16 rounding of zero flip their mask between the fp32 and the
            # bf16 forward and move single entries by a whole dy: bound the FRACTION of entries that are off instead
            if name in ("dx", "dres"):
                d = (got.float() - want).abs()
                assert float((d > 0.05 * want.abs().max()).float().mean()) < 5e-3, name
            else:  # reductions over all rows: the flipped entries average out
                assert rel(got.float(), want) < 0.2, name


def test_backbone_forward_train_matches_module_path(ME):
    """Res16UNet34C.forward_train (fused bf16 autograd units) against the fp32 module path: outputs and parameter gradients."""
    from room_instance_segmentation_mask3d_b200 import mask3d_model as MM, weights

    coords = torch.from_numpy(_batch(23)).cuda()
    feats = torch.ones(coords.shape[0], 1, device="cuda")
    model = MM.Res16UNet34C()
    weights.load_synth(model, seed=2)
    model = model.cuda().train()
    g = torch.Generator(device="cuda").manual_seed(5)

    def run(path, mode):
        ME.set_compute_mode(mode)
        for p in model.parameters():
            p.grad = None
        out, fm = path(ME.SparseTensor(feats, coordinates=coords))
        # quadratic loss on every returned level: a random LINEAR functional makes the norm-parameter gradients sums of
        # randomly signed terms, i.e. pure cancellation, and then bf16 rounding noise is as large as the signal
        tgt = [torch.randn(t.F.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7 + i)) for i, t in enumerate([out] + fm[:-1])]
        loss = sum((t.F.float() - 0.5 * d).square().mean() for t, d in zip([out] + fm[:-1], tgt))
        loss.backward()
        return out.F.detach().float(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    try:
        y_ref, g_ref = run(model.forward, "fp32")
        y_mod, g_mod = run(model.forward, "bf16")
        y_new, g_new = run(model.forward_train, "bf16")
    finally:
        ME.set_compute_mode("bf16")
    import numpy as np

    assert rel(y_new, y_ref) < 6e-2
    assert set(g_new) == set(g_ref)
    # Yardstick = the bf16 MODULE path (fp32 activations, bf16 conv operands): on a 4 k-voxel scene with random weights the
    # 60-layer instance-normalised network amplifies operand rounding so much that even it only reaches a median kernel-gradient
    # cosine of ~0.96 against fp32 (0.99 at 96 k voxels, scripts/grad_agreement.py).  The fused path (bf16 activations and
    # gradients as well) must stay within a small margin of that and far away from "unrelated".
    kern = [n for n in g_ref if n.endswith("kernel")]
    c_new = np.array([_cos(g_new[n], g_ref[n]) for n in kern])
    c_mod = np.array([_cos(g_mod[n], g_ref[n]) for n in kern])
    assert c_new.min() > 0.7, f"worst kernel-gradient cosine {c_new.min()} ({kern[int(c_new.argmin())]})"
    assert np.median(c_new) > np.median(c_mod) - 0.05 and np.median(c_new) > 0.9
    norm_names = [n for n in g_ref if not n.endswith("kernel")]
    cat = lambda g: torch.cat([g[n].flatten().float() for n in norm_names])
    assert _cos(cat(g_new), cat(g_ref)) > _cos(cat(g_mod), cat(g_ref)) - 0.1 and _cos(cat(g_new), cat(g_ref)) > 0.75
