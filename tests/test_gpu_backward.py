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
