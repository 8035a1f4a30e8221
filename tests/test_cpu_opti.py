"""CPU: host logic of BASELINE config 3 (LightGCNOpti / SpreadLightGCNOpti) against vectors recorded from the
reference's own model/LightGCNOpti/model.py (tests/golden/opti_tiny.npz, oracle/make_golden.py)."""
import io
import os

import numpy as np
import pandas as pd
import torch

import _stub_const

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _feature_frames(z):
    """user_features.csv / item_features.csv as main.py:38-40 reads them: tab-separated, list-literal strings,
    rows NOT sorted by id (Appendix B of SURVEY.md)."""
    ufd = pd.DataFrame({"user_id": z["perm_u"], "user_features": [str(z["user_feat"][u].tolist()) for u in z["perm_u"]]})
    ifd = pd.DataFrame({"item_id": z["perm_i"], "item_features": [str(z["item_feat"][i].tolist()) for i in z["perm_i"]]})
    return ufd, ifd


def test_opti_constructor_matches_reference_bit_for_bit():
    """torch.manual_seed(seed) -> LightGCNOpti(...): the two nn.Linear layers consume the RNG in the reference's
    order, so e^0 = Linear(features) equals the reference's initial weights exactly (model.py:36-49)."""
    _stub_const.install(model="LightGCNOpti")
    from model.LightGCNOpti.model import LightGCNOpti

    z = np.load(os.path.join(G, "opti_tiny.npz"))
    torch.manual_seed(42)
    m = LightGCNOpti(96, 160, 64, 3, torch.from_numpy(z["user_feat"]).float(), torch.from_numpy(z["item_feat"]).float())
    assert np.array_equal(m.users_emb.weight.detach().numpy(), z["users_w"])
    assert np.array_equal(m.items_emb.weight.detach().numpy(), z["items_w"])
    assert m.users_emb.weight.requires_grad and m.users_emb.weight.is_leaf       # a trainable Parameter, not a view
    assert {n for n, _ in m.named_parameters()} == {"user_linear.weight", "user_linear.bias", "item_linear.weight",
                                                    "item_linear.bias", "users_emb.weight", "items_emb.weight"}
    buf = io.BytesIO()
    torch.save(m, buf)                          # the reference pickles the whole module (train.py:189)
    buf.seek(0)
    m2 = torch.load(buf, weights_only=False)
    assert torch.equal(m2.items_emb.weight, m.items_emb.weight) and m2.layers == 3


def test_parse_features_tab_separated_literals(tmp_path):
    """The feature CSVs are TAB separated with python-list literal strings (main.py:38-40; recommend.py:150-163)."""
    _stub_const.install(model="LightGCNOpti")
    from model.LightGCNOpti.recommend import parse_features

    z = np.load(os.path.join(G, "opti_tiny.npz"))
    ufd, ifd = _feature_frames(z)
    ufd.to_csv(tmp_path / "user_features.csv", sep="\t", index=False)
    ifd.to_csv(tmp_path / "item_features.csv", sep="\t", index=False)
    u2 = pd.read_csv(tmp_path / "user_features.csv", sep="\t")
    i2 = pd.read_csv(tmp_path / "item_features.csv", sep="\t")
    assert isinstance(u2["user_features"].iloc[0], str)
    uf = parse_features(u2, "user_id", "user_features")
    itf = parse_features(i2, "item_id", "item_features")
    assert uf.dtype == torch.float32 and uf.shape == (96, 29) and itf.shape == (160, 31)
    assert torch.equal(uf, torch.from_numpy(z["user_feat"]).float())            # rows re-ordered by id
    assert torch.equal(itf, torch.from_numpy(z["item_feat"]).float())
    # already-parsed lists (a DataFrame built in memory) go through unchanged
    u3 = pd.DataFrame({"user_id": z["perm_u"], "user_features": [z["user_feat"][u].tolist() for u in z["perm_u"]]})
    assert torch.equal(parse_features(u3, "user_id", "user_features"), uf)


def test_reference_export_names_of_config3():
    _stub_const.install(model="SpreadLightGCNOpti")
    import model.SpreadLightGCNOpti.model as M
    import model.SpreadLightGCNOpti.recommend as R

    for name in ("getLightGCNOptiModel", "getAllocateMat", "getHybridSResourceMat", "getResourceMat"):
        assert callable(getattr(M, name)), name      # /root/reference/model/SpreadLightGCNOpti/model.py:25,98,173,191
    for name in ("recommendForAllUser", "recommendSpreadLightGCNOpti"):
        assert callable(getattr(R, name)), name      # recommend.py:18,56
