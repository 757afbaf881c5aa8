"""GPU parity of the whole hot path against the golden fixture written by the reference's UNMODIFIED model code
(tests/golden/make_golden.py: reference mask3d/models/{res16unet,mask3d}.py on the CPU oracle of MinkowskiEngine).

fp32 mode: <= 1e-3 relative (BASELINE.json north_star).  bf16 tensor-core mode: stated tolerance 5e-2 relative on the
backbone features (62 stacked bf16 convolutions, each re-normalised by instance norm)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "mask3d_tiny2.npz")


def rel(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / b.abs().max())


@pytest.fixture(scope="module")
def setup():
    from room_instance_segmentation_mask3d_b200 import mask3d_model, me_api, ops, weights

    g = np.load(GOLD)
    model = mask3d_model.Mask3D().eval()
    weights.load_synth(model, seed=0)
    model = model.cuda()
    return g, model, me_api, ops


def _input(g, me_api):
    return me_api.SparseTensor(features=torch.from_numpy(g["feats"]).cuda(), coordinates=torch.from_numpy(g["coords"]).cuda())


def test_level_coordinates_and_maps_bit_exact(setup):
    g, model, ME, ops = setup
    x = _input(g, ME)
    mgr = x.coordinate_manager
    key = x.coordinate_map_key
    keys = [key]
    for lvl in range(1, 5):
        keys.append(mgr.stride(keys[-1], 2))
    for lvl, k in enumerate(keys):
        assert np.array_equal(mgr.get_coordinates(k).cpu().numpy(), g[f"coords_l{lvl}"]), f"level {lvl}"
    assert np.array_equal(mgr.neighbor_table(keys[0], keys[0], 3, 1).nbr.cpu().numpy(), g["nbr3_l0"])
    assert np.array_equal(mgr.neighbor_table(keys[0], keys[1], 2, 1).nbr.cpu().numpy(), g["nbr2_l0_l1"])
    assert np.array_equal(mgr.neighbor_table(keys[0], keys[0], 5, 1).nbr.cpu().numpy().astype(np.int64).sum(1), g["nbr5_l0_rowsum"])


def test_backbone_fp32_vs_reference_golden(setup):
    g, model, ME, ops = setup
    ops.set_compute_mode("fp32")
    out, fmaps = model.backbone.forward_fused(_input(g, ME))
    assert rel(out.F, g["backbone_out"]) < 1e-3
    assert rel(fmaps[0].F, g["fmap0"]) < 1e-3
    assert rel(fmaps[3].F, g["fmap3"]) < 1e-3
    for i, fm in enumerate(fmaps):
        assert abs(float(fm.F.double().abs().sum()) - g["fmap_sums"][i, 1]) / g["fmap_sums"][i, 1] < 1e-4


def test_backbone_module_path_matches_fused(setup):
    """The ME-compatible modules composed like the reference (what unmodified reference code executes on the shim)."""
    g, model, ME, ops = setup
    ops.set_compute_mode("fp32")
    with torch.no_grad():
        out, fmaps = model.backbone(_input(g, ME))
    assert rel(out.F, g["backbone_out"]) < 1e-3
    assert rel(fmaps[3].F, g["fmap3"]) < 1e-3


def test_mask3d_fp32_vs_reference_golden(setup):
    g, model, ME, ops = setup
    ops.set_compute_mode("fp32")
    o = model(_input(g, ME))
    assert np.array_equal(o["sampled_coords"].cpu().numpy(), g["sampled_coords"]), "FPS picked different queries"
    pc = torch.stack(o["pred_class_all_decoders"]).cpu()
    assert rel(pc, g["pred_class"]) < 1e-3
    assert rel(o["pred_mask_all_decoders"][0], g["pred_mask_first"]) < 1e-3
    assert rel(o["pred_mask_all_decoders"][6], g["pred_mask_mid"]) < 1e-3
    assert rel(torch.cat(o["pred_mask"]), g["pred_mask_last"]) < 1e-3


def test_mask3d_bf16_tensor_core_mode(setup):
    g, model, ME, ops = setup
    ops.set_compute_mode("bf16")
    try:
        out, fmaps = model.backbone.forward_fused(_input(g, ME))
        e_bb = rel(out.F.float(), g["backbone_out"])
        o = model(_input(g, ME))
        e_mask = rel(torch.cat(o["pred_mask"]), g["pred_mask_last"])
        e_first = rel(o["pred_mask_all_decoders"][0], g["pred_mask_first"])
        print(f"bf16 mode: backbone rel err {e_bb:.3e}, first mask logits {e_first:.3e}, final mask logits {e_mask:.3e}")
        assert np.array_equal(o["sampled_coords"].cpu().numpy(), g["sampled_coords"])
        assert e_bb < 5e-2
        assert e_first < 5e-2
        # the final logits go through 12 data-dependent hard thresholds (sigmoid < 0.5): only sanity-bound them
        assert e_mask < 0.5
    finally:
        ops.set_compute_mode("bf16")


def test_mask3d_tf32_tensor_core_mode_vs_reference_golden(setup):
    """The reference's own precision (fp32 tensors, TF32 matmuls: main_instance_segmentation.py:52) on the tcgen05 kernel
    (tcgen05.mma.kind::tf32) against the fixture the reference's unmodified model wrote on the fp32 CPU oracle."""
    g, model, ME, ops = setup
    ops.set_compute_mode("tf32")
    try:
        out, fmaps = model.backbone.forward_fused(_input(g, ME))
        e_bb = rel(out.F.float(), g["backbone_out"])
        o = model(_input(g, ME))
        e_first = rel(o["pred_mask_all_decoders"][0], g["pred_mask_first"])
        e_mid = rel(o["pred_mask_all_decoders"][6], g["pred_mask_mid"])
        e_last = rel(torch.cat(o["pred_mask"]), g["pred_mask_last"])
        e_cls = rel(torch.stack(o["pred_class_all_decoders"]).cpu(), g["pred_class"])
        print(f"tf32 mode: backbone {e_bb:.3e}, mask logits first {e_first:.3e} mid {e_mid:.3e} last {e_last:.3e}, class {e_cls:.3e}")
        assert np.array_equal(o["sampled_coords"].cpu().numpy(), g["sampled_coords"])
        # one conv is <= 1e-3 (test_conv_tf32_tensor_core_parity); 62 instance-normalised layers in sequence accumulate TF32
        # operand rounding (2^-11 per operand) to a few 1e-3 on the backbone output, and the decoder's hard thresholds
        # (sigmoid < 0.5) can flip single voxels, so the later mask logits are bounded a little wider
        assert e_bb < 6e-3 and e_first < 6e-3   # measured 3.9e-3 / 2.6e-3
        assert e_mid < 3e-2                     # measured 7.6e-3
        # after 12 hard-thresholded layers single flipped voxels move a query: bound the binary masks, not the logits
        want = torch.from_numpy(g["pred_mask_last"]) > 0
        got = torch.cat(o["pred_mask"]).cpu() > 0
        flipped = float((want != got).float().mean())
        print(f"tf32 mode: flipped final mask bits {flipped:.4%}")
        assert flipped < 0.02 and e_last < 0.3
    finally:
        ops.set_compute_mode("bf16")


def test_bf16_mode_is_run_to_run_bit_identical_when_deterministic(setup):
    """mk_set_deterministic(1): two forwards of the same batch give bit-identical outputs in the benchmarked bf16 tensor-core mode
    (the reference trains with `deterministic: true`, conf/trainer/trainer.yaml:1); also with the next scene prefetched in between."""
    g, model, ME, ops = setup
    ops.set_compute_mode("bf16")
    ops.set_deterministic(True)
    try:
        runs = []
        for _ in range(3):
            o = model(_input(g, ME))
            torch.cuda.synchronize()
            runs.append([t.clone() for t in o["pred_mask_all_decoders"]] + [t.clone() for t in o["pred_class_all_decoders"]] +
                        [o["backbone_features"].F.clone()])
        for r in runs[1:]:
            for a, b in zip(runs[0], r):
                assert torch.equal(a, b)
    finally:
        ops.set_deterministic(False)


def test_final_prediction_only_forward_matches_the_full_forward(setup):
    """forward(aux_outputs=False) skips the intermediate outputs' full-resolution mask logits and nothing else: class logits of all 13
    outputs, the final mask logits and the queries are bit-identical to the full forward (deterministic mode, bf16 and fp32)."""
    g, model, ME, ops = setup
    ops.set_deterministic(True)
    try:
        for mode in ("bf16", "fp32"):
            ops.set_compute_mode(mode)
            full = model(_input(g, ME))
            lean = model(_input(g, ME), aux_outputs=False)
            torch.cuda.synchronize()
            assert len(full["pred_mask_all_decoders"]) == 13 and len(lean["pred_mask_all_decoders"]) == 1
            assert torch.equal(full["pred_mask_all_decoders"][-1], lean["pred_mask_all_decoders"][-1])
            assert len(lean["pred_class_all_decoders"]) == 13
            for a, b in zip(full["pred_class_all_decoders"], lean["pred_class_all_decoders"]):
                assert torch.equal(a, b)
            assert torch.equal(full["queries"], lean["queries"])
    finally:
        ops.set_deterministic(False)
        ops.set_compute_mode("fp32")


def test_prefetch_on_side_stream_gives_identical_results(setup):
    """Mask3D.forward(x, prefetch=next): the coordinate phase of the next scene runs on a side stream behind the current
    scene's backbone.  Outputs must equal the plain sequential calls bit for bit (fp32 mode), over several alternating scenes."""
    import numpy as np
    from room_instance_segmentation_mask3d_b200 import mask3d_model, me_api, ops, scenes, weights

    ops.set_compute_mode("fp32")
    try:
        model = mask3d_model.Mask3D().eval()
        weights.load_synth(model, seed=0)
        model = model.cuda()
        batches = [scenes.collate([scenes.tiny_scene(3 + j), scenes.tiny_scene(7 + j, (20, 16, 10))]) for j in range(3)]
        dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()) for c, f in batches]
        plain = [model(me_api.SparseTensor(features=f, coordinates=c)) for c, f in dev]
        x = model.prepare(*dev[0])
        for i in range(6):
            o = model(x, prefetch=dev[(i + 1) % 3])
            x = o.pop("prefetched")
            ref = plain[i % 3]
            assert torch.equal(o["fps_idx"], ref["fps_idx"])
            assert torch.equal(o["pred_mask_all_decoders"][-1], ref["pred_mask_all_decoders"][-1])
            assert torch.equal(o["pred_class"], ref["pred_class"])
        torch.cuda.synchronize()
    finally:
        ops.set_compute_mode("bf16")


def test_two_lane_alternation_gives_identical_results(setup):
    """Consecutive scenes alternating between two CUDA streams with the next scene prepared on the side stream (what bench.py
    does): every scene's outputs must equal the plain single-stream call bit for bit (fp32 mode), also under memory reuse
    across many steps."""
    from room_instance_segmentation_mask3d_b200 import mask3d_model, me_api, ops, scenes, weights

    ops.set_compute_mode("fp32")
    try:
        model = mask3d_model.Mask3D().eval()
        weights.load_synth(model, seed=0)
        model = model.cuda()
        batches = [scenes.collate([scenes.tiny_scene(11 + j), scenes.tiny_scene(17 + j, (20, 16, 10))]) for j in range(3)]
        dev = [(torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()) for c, f in batches]
        plain = [model(me_api.SparseTensor(features=f, coordinates=c)) for c, f in dev]
        want = [(p["fps_idx"].clone(), p["pred_mask_all_decoders"][-1].clone(), p["pred_class"].clone()) for p in plain]
        del plain
        lanes = [torch.cuda.Stream(), torch.cuda.Stream()]
        x = model.prepare(*dev[0])
        got = []
        for i in range(12):
            with torch.cuda.stream(lanes[i % 2]):
                o = model(x, prefetch=dev[(i + 1) % 3])
                x = o.pop("prefetched")
                got.append((i % 3, o["fps_idx"], o["pred_mask_all_decoders"][-1], o["pred_class"]))
            del o
        torch.cuda.synchronize()
        for j, a, b, c in got:
            assert torch.equal(a, want[j][0]) and torch.equal(b, want[j][1]) and torch.equal(c, want[j][2])
    finally:
        ops.set_compute_mode("bf16")


def test_mask3d_forward_train_gradients_vs_oracle(setup):
    """Whole-model training forward + backward (Mask3D.forward_train: tcgen05 conv fwd/dgrad/wgrad, mask-logit and masked
    cross-attention backward kernels, bf16 activations) against autograd through the fp32 CPU oracle of the reference model
    (oracle/mask3d_oracle.py, pinned by the reference's golden fixture) on the same weights, input and loss.
    Outputs: the tolerance of the bf16 inference test.  Gradients: cosine similarity per parameter — bf16 operand rounding
    through 62 instance-normalised convolutions does not allow an element-wise bound (scripts/grad_agreement.py).  Stated bars:
    decoder parameters (everything the new backward kernels feed directly) median >= 0.99, min >= 0.95 (measured 0.998 /
    0.983); backbone kernels median >= 0.85, min >= 0.75 (measured 0.89 / 0.82 on this 2 x 938-voxel fixture, whose coarse
    levels hold a handful of voxels per scene, so instance norm amplifies the bf16 rounding; the backbone's own backward is
    bounded against the fp32 module path in test_backbone_forward_train_matches_module_path)."""
    from oracle import mask3d_oracle, me_oracle
    from room_instance_segmentation_mask3d_b200 import weights

    g, model, ME, ops = setup
    ops.set_compute_mode("bf16")
    gen = torch.Generator().manual_seed(5)
    n = g["coords"].shape[0]
    D_out = 13
    Rm = [torch.randn(n, 25, generator=gen) for _ in range(D_out)]
    Rc = [torch.randn(2, 25, 2, generator=gen) for _ in range(D_out)]

    def loss_of(pred_masks, pred_class, dev):
        tot = 0
        for d in range(D_out):
            pm = pred_masks[d] if isinstance(pred_masks[d], torch.Tensor) else torch.cat(pred_masks[d])
            tot = tot + (pm.float() * Rm[d].to(dev)).mean() + 0.1 * pm.float().square().mean() + (pred_class[d].float() * Rc[d].to(dev)).sum()
        return tot

    model.train()
    for p_ in model.parameters():
        p_.grad = None
    o = model.forward_train(_input(g, ME))
    assert rel(o["pred_mask_all_decoders"][0], g["pred_mask_first"]) < 5e-2
    assert np.array_equal(o["sampled_coords"].detach().cpu().numpy(), g["sampled_coords"])
    loss_of(o["pred_mask_all_decoders"], o["pred_class_all_decoders"], "cuda").backward()
    got = {k: v.grad.detach().cpu() for k, v in model.named_parameters() if v.grad is not None}
    params = {k for k, _ in model.named_parameters()}
    model.eval()

    spec = [(nm, tuple(int(d) for d in s.split(",") if d)) for nm, s in zip(g["names"], g["shapes"])]
    sd = {k: torch.from_numpy(v) for k, v in weights.synth_state_dict(spec).items()}
    for v in sd.values():
        if v.is_floating_point():
            v.requires_grad_(True)
    x = me_oracle.SparseTensor(features=torch.from_numpy(g["feats"]), coordinates=torch.from_numpy(g["coords"]))
    ref = mask3d_oracle.mask3d_forward(sd, x)
    loss_of(ref["pred_masks"], ref["pred_class"], "cpu").backward()

    cos = {}
    for k, v in sd.items():
        if v.grad is None or k.startswith("backbone.final") or k not in params:  # buffers (pos_enc.gauss_B) are not trained
            continue
        assert k in got, f"no gradient reached {k}"
        a, b = got[k].double().flatten(), v.grad.double().flatten()
        assert torch.isfinite(a).all(), k
        cos[k] = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
    dec = sorted(c for k, c in cos.items() if not k.startswith("backbone."))
    bb = sorted(c for k, c in cos.items() if k.startswith("backbone.") and k.endswith("kernel"))
    worst = sorted(cos.items(), key=lambda kv: kv[1])[:5]
    print(f"forward_train vs fp32 oracle: decoder cos median {dec[len(dec) // 2]:.4f} min {dec[0]:.4f} ({len(dec)} tensors); "
          f"backbone kernels median {bb[len(bb) // 2]:.4f} min {bb[0]:.4f} ({len(bb)}); worst {worst}")
    assert dec[len(dec) // 2] >= 0.99 and dec[0] >= 0.95
    assert bb[len(bb) // 2] >= 0.85 and bb[0] >= 0.75


@pytest.mark.timeout(900)
def test_bf16_tensor_core_mode_binary_masks_vs_fp32_oracle_on_a_room_scene():
    """Parity of the BENCHMARKED mode (bf16 tensor cores) where it matters: the binary instance masks.  One S200k-class room
    (223 648 voxels at 2 cm) through the fused bf16 path against the fp32 CPU oracle of the reference's
    algorithm: FPS queries identical, per-query binary-mask IoU and the fraction of flipped (voxel, query) bits of the final
    output and of the first decoder output (which has no data-dependent threshold upstream), tf32 mode beside it."""
    from oracle import mask3d_oracle, me_oracle
    from room_instance_segmentation_mask3d_b200 import mask3d_model, me_api, ops, scenes, weights

    coords, feats = scenes.collate([scenes.box_room()])
    model = mask3d_model.Mask3D().eval()
    weights.load_synth(model, seed=0)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        ref = mask3d_oracle.mask3d_forward(sd, me_oracle.SparseTensor(features=torch.from_numpy(feats), coordinates=torch.from_numpy(coords)))
    model = model.cuda()

    def stats(got, want):
        g, w = got > 0, want > 0
        flipped = float((g != w).float().mean())
        inter, union = (g & w).sum(0).float(), (g | w).sum(0).float()
        iou = torch.where(union > 0, inter / union.clamp_min(1), torch.ones_like(union))
        return flipped, float(iou.min()), float(iou.mean())

    res = {}
    try:
        for mode in ("bf16", "tf32"):
            ops.set_compute_mode(mode)
            o = model(me_api.SparseTensor(features=torch.from_numpy(feats).cuda(), coordinates=torch.from_numpy(coords).cuda()))
            assert np.array_equal(o["sampled_coords"].cpu().numpy(), ref["sampled_coords"].numpy()), "FPS picked different queries"
            first = stats(o["pred_mask_all_decoders"][0].float().cpu(), torch.cat(ref["pred_masks"][0]))
            last = stats(o["pred_mask_all_decoders"][-1].float().cpu(), torch.cat(ref["pred_masks"][-1]))
            res[mode] = (first, last)
            print(f"{mode}: first output flipped {first[0]:.4%} IoU min {first[1]:.4f} mean {first[2]:.4f} | final output flipped "
                  f"{last[0]:.4%} IoU min {last[1]:.4f} mean {last[2]:.4f}")
    finally:
        ops.set_compute_mode("bf16")
    # Measured on B200 (synthetic, untrained weights: mask logits hover around 0, the worst case for sign agreement):
    #   first decoder output (no data-dependent threshold upstream)   bf16: 0.44 % of the bits flipped, IoU 0.940;  tf32: 0.04 %, IoU 0.994
    #   final output (12 hard-thresholded attention layers upstream)  bf16: 4.5 % flipped, mean IoU 0.81;         tf32: 1.5 %, mean IoU 0.93
    # tf32 is the reference's OWN precision (main_instance_segmentation.py:52): its final masks move by the same order of magnitude
    # against an fp32 evaluation, i.e. the spread of the final output is a property of the thresholded decoder, not of bf16.
    for mode, (flip1, iou1, flipL, iouL) in (("bf16", (0.01, 0.90, 0.08, 0.70)), ("tf32", (0.002, 0.98, 0.03, 0.85))):
        first, last = res[mode]
        assert first[0] <= flip1 and first[2] >= iou1, (mode, "first decoder output", first)
        assert last[0] <= flipL and last[2] >= iouL, (mode, "final output", last)
