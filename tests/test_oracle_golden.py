"""CPU: the oracle restatement and the host-side drop-ins against golden vectors produced by the LIVE reference
(oracle/make_golden.py). No reference checkout and no GPU is needed to run these."""
import json
import os

import pytest
import torch

from helpers import FULL_CFG, TINY_CFG, build_ours, checksums
from oracle import restatement as R


@pytest.fixture(scope="module")
def tiny(golden_dir):
    return torch.load(os.path.join(golden_dir, "stack1_tiny.pt"), weights_only=False)


@pytest.fixture(scope="module")
def host(golden_dir):
    return torch.load(os.path.join(golden_dir, "host_fns.pt"), weights_only=False)


def test_init_is_rng_identical_to_reference(golden_dir):
    """Our constructors consume the torch RNG exactly like the reference's: same seed -> same 289 tensors."""
    gold = json.load(open(os.path.join(golden_dir, "init_checksums.json")))
    model = build_ours(FULL_CFG, seed=gold["seed"])
    sd = model.state_dict()
    assert len(sd) == gold["num_keys"] == 289
    assert list(sd.keys()) == list(gold["checksums"].keys())
    mine = checksums(sd)
    for k, (s, a) in gold["checksums"].items():
        assert mine[k][0] == pytest.approx(s, rel=1e-9, abs=1e-9), k
        assert mine[k][1] == pytest.approx(a, rel=1e-9, abs=1e-9), k


def _tiny_state(tiny):
    model = build_ours(TINY_CFG, seed=tiny["seed"])
    sd = R.perturb_state_dict(model.state_dict(), seed=tiny["perturb_seed"])
    mine = checksums(sd)
    for k, (s, a) in tiny["state_checksums"].items():
        assert mine[k][1] == pytest.approx(a, rel=1e-9, abs=1e-9), k
    return sd


def test_restatement_matches_reference_forward_and_backward(tiny):
    sd = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and "running" not in k else v) for k, v in _tiny_state(tiny).items()}
    m = TINY_CFG["model"]
    pred = R.tiny_video_mae_forward(sd, tiny["clip"], tiny["mask"], m["decoder_depth"], m["decoder_num_heads"])
    assert torch.allclose(pred, tiny["pred"], rtol=1e-4, atol=2e-5)
    loss = R.masked_mse(pred, tiny["clip"], tiny["mask"], 1, 8, True)
    assert float(loss) == pytest.approx(float(tiny["loss"]), rel=1e-5)
    loss.backward()
    for k, gref in tiny["grads"].items():
        g = sd[k].grad
        assert g is not None, k
        assert torch.allclose(g, gref, rtol=2e-3, atol=1e-6 + 1e-4 * float(gref.abs().max())), k
    for k, n in tiny["grad_norms"].items():
        if n > 1e-6 and sd[k].grad is not None:
            assert float(sd[k].grad.norm()) == pytest.approx(n, rel=5e-3), k


def test_tube_mask_bit_exact(host):
    from ssl_vit_video_analytics_b200.datasets.mae_loader import get_tube_mask, visible_indices

    for fn in (get_tube_mask, R.get_tube_mask):
        torch.manual_seed(42)
        assert torch.equal(fn(4, 3, 196, 0.9), host["tube_mask_b4_t3_l196_r0.9_seed42"])
        torch.manual_seed(5)
        assert torch.equal(fn(2, 8, 49, 0.5), host["tube_mask_b2_t8_l49_r0.5_seed5"])
    m = host["tube_mask_b4_t3_l196_r0.9_seed42"]
    assert m.dtype == torch.bool and m.shape == (4, 3, 196) and int(m[0, 0].sum()) == int(0.9 * 196)
    vis = visible_indices(m)
    assert vis.shape == (4, 196 - 176) and vis.dtype == torch.int32
    for b in range(4):
        assert torch.equal(vis[b].long(), torch.nonzero(~m[b, 0]).flatten())


def test_token_masks_bit_exact(host):
    from ssl_vit_video_analytics_b200.mae import masking

    for mod in (masking, R):
        torch.manual_seed(11)
        assert torch.equal(mod.make_token_mask(3, 4, 9, 0.6, "tube", torch.device("cpu")), host["token_mask_tube_b3_t4_p9_r0.6_seed11"])
        torch.manual_seed(12)
        assert torch.equal(mod.make_token_mask(3, 4, 9, 0.6, "random", torch.device("cpu")), host["token_mask_random_b3_t4_p9_r0.6_seed12"])
    sched = [{"start": 1, "end": 5, "value": 0.5}, {"start": 5, "end": 9, "value": 0.75}]
    assert [masking.get_mask_ratio(e, sched, 0.9) for e in range(0, 12)] == host["mask_ratio_schedule"]
    with pytest.raises(ValueError):
        masking.make_token_mask(1, 1, 4, 0.5, "bogus", torch.device("cpu"))
    m = host["token_mask_tube_b3_t4_p9_r0.6_seed11"]
    assert masking.count_masked(m) + masking.count_visible(m) == m.numel()


def test_patchify_layout(host):
    from ssl_vit_video_analytics_b200.train_ssl_mae import patchify, unpatchify

    imgs = torch.arange(2 * 3 * 2 * 16 * 16, dtype=torch.float32).reshape(2, 3, 2, 16, 16)
    gold = host["patchify_arange_2x3x2x16x16_p8"]
    assert torch.equal(patchify(imgs, 8), gold)
    assert torch.equal(R.patchify(imgs, 1, 8), gold)
    assert torch.equal(unpatchify(gold, 2, 16, 16, 8), imgs)


def test_stack2_losses(host):
    """The oracle's losses against the live-reference golden values; ours are CUDA-only (checked against the same goldens in
    tests/test_kernels_gpu.py) and must refuse CPU tensors."""
    from ssl_vit_video_analytics_b200.mae import losses

    p, t = host["loss_inputs"]
    assert torch.allclose(R.mae_l2_loss(p, t, True), host["mae_l2_norm"], rtol=1e-6)
    assert torch.allclose(R.mae_l2_loss(p, t, False), host["mae_l2_raw"], rtol=1e-6)
    assert torch.allclose(R.mae_l1_loss(p, t), host["mae_l1"], rtol=1e-6)
    assert torch.allclose(R.mae_cosine_loss(p, t), host["mae_cos"], rtol=1e-6)
    assert torch.allclose(R.mae_recon_loss(p, t, "l1", False), host["mae_recon_positional"], rtol=1e-6)
    stats = losses.reconstruction_error_stats(p, t)
    for k, v in host["recon_stats"].items():
        assert stats[k] == pytest.approx(v, rel=1e-6)
    with pytest.raises(RuntimeError, match="no CPU path"):
        losses.mae_l2_loss(p, t, True)
    with pytest.raises(RuntimeError, match="no CPU path"):
        losses.mae_recon_loss(p, t, "cosine", False)


def test_fedavg_known_answer(host):
    import torch.nn as nn

    from ssl_vit_video_analytics_b200.federated.fed_loop import fedavg_aggregate

    glob = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4))
    glob.load_state_dict(host["fedavg_global_before"])
    new = fedavg_aggregate(glob, host["fedavg_clients"], host["fedavg_weights"])
    ref = host["fedavg_result"]
    assert list(new.keys()) == list(ref.keys())
    for k in ref:
        assert torch.equal(new[k], ref[k]), k            # same summation order -> bit-exact
        assert torch.equal(glob.state_dict()[k], ref[k]), k
    ora = R.fedavg(host["fedavg_global_before"], host["fedavg_clients"], host["fedavg_weights"])
    for k in ref:
        assert torch.equal(ora[k], ref[k]), k
    assert int(new["1.num_batches_tracked"]) == 4           # max over clients (2, 3, 4)
    with pytest.raises(RuntimeError):
        fedavg_aggregate(glob, [], [])
    with pytest.raises(RuntimeError):
        fedavg_aggregate(glob, host["fedavg_clients"], [1.0])
    with pytest.raises(RuntimeError):
        fedavg_aggregate(glob, host["fedavg_clients"], [0.0, 0.0, 0.0])
    # a key missing from one client keeps the global value
    partial = [dict(c) for c in host["fedavg_clients"]]
    del partial[1]["0.bias"]
    before = glob.state_dict()["0.bias"].clone()
    out = fedavg_aggregate(glob, partial, host["fedavg_weights"])
    assert torch.equal(out["0.bias"], before)


def test_noise_identity_for_nonpositive_sigma(host):
    from ssl_vit_video_analytics_b200.privacy.feature_noise import add_gaussian_noise

    assert host["noise_sigma0_is_identity"] is True
    z = torch.randn(3, 5)
    assert add_gaussian_noise(z, 0.0) is z and add_gaussian_noise(z, -1.0) is z
    with pytest.raises(RuntimeError):
        add_gaussian_noise(z, 0.1)   # CPU tensor: there is no CPU path


# ---- round 2: the reference's literal defaults (checkpointing + drop-path) and the stack-2 mask/gather logic -----------
@pytest.fixture(scope="module")
def defaults(golden_dir):
    return torch.load(os.path.join(golden_dir, "stack1_defaults.pt"), weights_only=False)


@pytest.fixture(scope="module")
def stack2(golden_dir):
    return torch.load(os.path.join(golden_dir, "stack2_mae.pt"), weights_only=False)


def test_restatement_with_recorded_drop_path_masks_matches_reference_defaults(defaults):
    """Live reference with use_checkpoint=True, drop_path_rate=0.1 (its defaults) vs the oracle fed the recorded keep masks:
    checkpointing must not change pred / loss / gradients, drop-path enters only through the masks."""
    d = defaults
    model = build_ours(d["config"], seed=d["seed"])
    sd0 = R.perturb_state_dict(model.state_dict(), seed=d["perturb_seed"])
    sd = {k: (v.clone().requires_grad_(True) if torch.is_floating_point(v) and "running" not in k else v) for k, v in sd0.items()}
    m = d["config"]["model"]
    assert sorted(d["drop_masks"]) == ["stages.0.1", "stages.1.1", "stages.1.2"] + [f"stages.2.{j}" for j in range(1, 7)]
    assert any(float(t.min()) == 0.0 for v in d["drop_masks"].values() for t in v[1:]), "the golden must actually drop a path"
    pred = R.tiny_video_mae_forward(sd, d["clip"], d["mask"], m["decoder_depth"], m["decoder_num_heads"], drop_masks=d["drop_masks"])
    assert torch.allclose(pred, d["pred"], rtol=1e-4, atol=3e-5)
    loss = R.masked_mse(pred, d["clip"], d["mask"], 1, 8, True)
    assert float(loss) == pytest.approx(float(d["loss"]), rel=1e-5)
    loss.backward()
    for k, gref in d["grads"].items():
        g = sd[k].grad
        assert g is not None, k
        assert torch.allclose(g, gref, rtol=2e-3, atol=1e-6 + 2e-4 * float(gref.abs().max())), k
    # BatchNorm layers inside checkpointed stages count (and update) twice per step, the un-checkpointed stem once
    nbt = d["num_batches_tracked"]
    assert nbt["encoder.patch_embed.patch_embed.0.bn.num_batches_tracked"] == 1
    assert nbt["encoder.stages.0.0.conv.0.bn.num_batches_tracked"] == 2
    assert nbt["encoder.stages.2.0.conv.5.bn.num_batches_tracked"] == 2
    assert nbt["encoder.stages.3.0.conv.0.bn.num_batches_tracked"] == 0      # stage 3 is not executed by the MAE


def test_stack2_restatement_matches_reference_mask_and_gather(stack2):
    s = stack2
    dec = s["decoder"]
    tok = s["tokens"].reshape(s["B"], s["T"] * s["P"], s["D"])
    for mode, case in s["cases"].items():
        sd = {k: v.clone().requires_grad_(True) for k, v in s["state"].items()}
        scale = sd["encoder.scale"]
        pred, target = R.tinyvit_mae_forward(sd, tok * scale, case["token_mask"], dec["depth"], dec["num_heads"])
        assert pred.shape == case["pred"].shape and torch.equal(target, case["target"]), mode       # gather rows bit-exact
        assert torch.allclose(pred, case["pred"], rtol=1e-4, atol=2e-5), mode
        loss = R.mae_l2_loss(pred, target, True)
        assert float(loss) == pytest.approx(float(case["loss"]), rel=1e-5)
        loss.backward()
        for k, gref in case["grads"].items():
            assert torch.allclose(sd[k].grad, gref, rtol=2e-3, atol=1e-7 + 2e-4 * float(gref.abs().max())), (mode, k)


def test_frame_window_sampling_matches_reference_semantics():
    """mae_loader.py:44-59 restated: always clip_len indices; short videos -> truncated linspace; long videos -> random start + stride
    drawn from numpy's global RNG exactly like the reference (same seed -> same window)."""
    import numpy as np

    from ssl_vit_video_analytics_b200.datasets.mae_loader import sample_frame_indices

    assert list(sample_frame_indices(10, 16, 2)) == list(np.linspace(0, 9, 16).astype(int))
    assert len(sample_frame_indices(1, 16, 2)) == 16 and set(sample_frame_indices(1, 16, 2)) == {0}
    np.random.seed(5)
    start = np.random.randint(0, 100 - 32 + 1)
    np.random.seed(5)
    got = sample_frame_indices(100, 16, 2)
    assert list(got) == list(range(start, start + 32, 2)) and len(got) == 16
    assert list(sample_frame_indices(32, 16, 2)) == list(range(0, 32, 2))
