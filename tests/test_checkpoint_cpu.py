"""CPU: checkpoint hand-off formats of the reference — bare encoder files, `encoder.` prefixes, wrappers,
the stack-2 dictionary. Key names are checked against the golden key list taken from the live reference."""
import json
import os

import torch

from helpers import TINY_CFG, build_ours
from ssl_vit_video_analytics_b200 import checkpoint as C

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "init_checksums.json")


def test_encoder_checkpoint_round_trip_and_reference_key_names(tmp_path):
    model = build_ours(TINY_CFG, seed=1)
    path = C.save_encoder_checkpoint(model, tmp_path / "ckpt", 10)
    assert path.name == "encoder_ep10.pth" and path.exists()
    saved = torch.load(path, map_location="cpu")
    ref_keys = set(json.load(open(GOLDEN))["checksums"].keys())
    assert {"encoder." + k for k in saved} == {k for k in ref_keys if k.startswith("encoder.")}      # bare encoder keys == reference's
    other = build_ours(TINY_CFG, seed=2)
    missing, unexpected = C.load_encoder_checkpoint(other.encoder, path)
    assert not missing and not unexpected
    for (k, a), (_, b) in zip(model.encoder.state_dict().items(), other.encoder.state_dict().items()):
        assert torch.equal(a, b), k


def test_prefixed_and_wrapped_formats(tmp_path):
    model = build_ours(TINY_CFG, seed=1)
    whole = model.state_dict()                                   # encoder.* + decoder keys, as a reference TinyVideoMAE checkpoint
    for obj in (whole, {"model": whole}, {"state_dict": whole}, {"net": C.add_prefix(model.encoder.state_dict())}):
        other = build_ours(TINY_CFG, seed=3)
        missing, unexpected = C.load_encoder_checkpoint(other.encoder, obj)
        assert not missing and not unexpected
        assert all(torch.equal(a, b) for a, b in zip(model.encoder.state_dict().values(), other.encoder.state_dict().values()))
    assert C.extract_state_dict({"epoch": 3}) is None and C.extract_state_dict(None) is None

    class Wrapper(torch.nn.Module):                              # the finetune model keeps the encoder under .backbone
        def __init__(self, enc):
            super().__init__()
            self.backbone = enc

    w = Wrapper(build_ours(TINY_CFG, seed=4).encoder)
    p = tmp_path / "ssl.pth"
    torch.save({"model": whole}, p)
    assert C.load_pretrained_ssl(w, str(p)) is True
    assert all(torch.equal(a, b) for a, b in zip(model.encoder.state_dict().values(), w.backbone.state_dict().values()))
    assert C.load_pretrained_ssl(w, str(tmp_path / "missing.pth")) is False


def test_stack2_checkpoint_dictionary(tmp_path):
    model = torch.nn.Linear(4, 3)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: 1.0)
    model(torch.randn(2, 4)).sum().backward()
    opt.step()
    cfg = {"b": 1, "a": {"z": 2}}
    p = tmp_path / "run" / "last.pth"
    C.save_mae_checkpoint(p, model, opt, None, sched, 7, {"loss": 0.5}, cfg)
    ck = torch.load(p, map_location="cpu")
    assert set(ck) == {"epoch", "model", "optimizer", "scaler", "scheduler", "best", "cfg_hash"} and ck["scaler"] is None
    assert ck["cfg_hash"] == C.config_hash({"a": {"z": 2}, "b": 1}) and len(ck["cfg_hash"]) == 12
    m2 = torch.nn.Linear(4, 3)
    o2 = torch.optim.AdamW(m2.parameters(), lr=1e-3)
    out = C.load_mae_checkpoint(p, m2, o2)
    assert out["epoch"] == 7 and torch.equal(m2.weight, model.weight)
