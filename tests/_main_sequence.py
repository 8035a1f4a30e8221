"""TEST INFRASTRUCTURE — the call sequence of the reference's main.py, restated so that it can run on the GPU box
(where /root/reference does not exist; where it does, `run_main.py` executes the real main.py on the drop-in).

    Step 1 (main.py:27-40)  read filter_rating / train / val / test CSVs and the TAB-separated feature CSVs from
                            cfg.PREPROCESSING["save_path"]; user_num / item_num = number of unique ids (:47-49)
    Step 2 (main.py:65-80)  dispatch on cfg.MODEL["name"] to one of the five recommend* entry points
    Step 3 (main.py:85-101) recommendDictToTensor, positive dicts, item degrees over train+val, dense interaction
                            matrix, getAccurateMetrics on the TEST positives, getDiversityMetrics
Only modules that main.py itself imports are used (main.py:12-23)."""
import numpy as np
import pandas as pd


def write_preprocessed(save_path: str, d, splits, user_feat=None, item_feat=None, seed: int = 0):
    """The on-disk inputs of main.py for a synthetic interaction set."""
    tr, va, te = splits
    rng = np.random.default_rng(seed)
    frame = lambda idx: pd.DataFrame({"user_id": d.users[idx], "item_id": d.items[idx], "rating": 5,  # noqa: E731
                                      "rating_time": "2024-01-01 00:00:00"})
    frame(np.arange(d.users.size)).to_csv(save_path + "filter_rating.csv", index=False)
    frame(tr).to_csv(save_path + "train_data.csv", index=False)
    frame(va).to_csv(save_path + "val_data.csv", index=False)
    frame(te).to_csv(save_path + "test_data.csv", index=False)
    if user_feat is None:
        user_feat = np.round(rng.random((d.n_users, 29)), 3)
        item_feat = np.round(rng.random((d.n_items, 31)), 3)
    pu, pi = rng.permutation(d.n_users), rng.permutation(d.n_items)
    pd.DataFrame({"user_id": pu, "user_features": [str(user_feat[u].tolist()) for u in pu]}).to_csv(
        save_path + "user_features.csv", sep="\t", index=False)
    pd.DataFrame({"item_id": pi, "item_features": [str(item_feat[i].tolist()) for i in pi]}).to_csv(
        save_path + "item_features.csv", sep="\t", index=False)
    return user_feat, item_feat


def run(cfg):
    from metrics.accurate import getAccurateMetrics
    from metrics.diversity import getDiversityMetrics
    from model.LightGCN.recommend import recommendLightGCN
    from model.LightGCNOpti.recommend import recommendLightGCNOpti
    from model.SpreadLightGCN.recommend import recommendSpreadLightGCN
    from model.SpreadLightGCNOpti.recommend import recommendSpreadLightGCNOpti
    from model.SpreadMethod.recommend import recommendSpreadMethod
    from utils.trans import (getInteractionMatrixByDataframe, getItemDegreeByUserPosItemDict,
                             getUserItemsDictByDataframe, recommendDictToTensor)

    sp = cfg.PREPROCESSING["save_path"]
    rating_df = pd.read_csv(sp + "filter_rating.csv")
    train_data_df = pd.read_csv(sp + "train_data.csv")
    val_data_df = pd.read_csv(sp + "val_data.csv")
    test_data_df = pd.read_csv(sp + "test_data.csv")
    user_features_df = pd.read_csv(sp + "user_features.csv", sep="\t")
    item_features_df = pd.read_csv(sp + "item_features.csv", sep="\t")
    user_num = len(rating_df["user_id"].unique())
    item_num = len(rating_df["item_id"].unique())
    name = cfg.MODEL["name"]
    if name in ("ProbS", "HeatS", "HybridS"):
        rec = recommendSpreadMethod(user_num, item_num, train_data_df, val_data_df, name)
    elif name == "LightGCN":
        rec = recommendLightGCN(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df)
    elif name == "LightGCNOpti":
        rec = recommendLightGCNOpti(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df,
                                    user_features_df, item_features_df)
    elif name == "SpreadLightGCN":
        rec = recommendSpreadLightGCN(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df)
    elif name == "SpreadLightGCNOpti":
        rec = recommendSpreadLightGCNOpti(user_num, item_num, rating_df, train_data_df, val_data_df, test_data_df,
                                          user_features_df, item_features_df)
    else:
        raise ValueError(name)
    recommendations = recommendDictToTensor(rec)
    train_pos = getUserItemsDictByDataframe(train_data_df)
    val_pos = getUserItemsDictByDataframe(val_data_df)
    test_pos = getUserItemsDictByDataframe(test_data_df)
    item_degree_dict = getItemDegreeByUserPosItemDict(train_pos, val_pos)
    interaction_mat = getInteractionMatrixByDataframe(user_num, item_num, pd.concat([train_data_df, val_data_df]))
    k = cfg.RECOMMEND["k"]
    acc = getAccurateMetrics(test_pos, recommendations, k)
    div = getDiversityMetrics(recommendations, item_degree_dict, interaction_mat, k)
    return {"rec": rec, "recommendations": recommendations, "accurate": acc, "diversity": div, "user_num": user_num,
            "item_num": item_num, "test_pos": test_pos, "item_degree_dict": item_degree_dict,
            "interaction_mat": interaction_mat, "frames": (rating_df, train_data_df, val_data_df, test_data_df)}
