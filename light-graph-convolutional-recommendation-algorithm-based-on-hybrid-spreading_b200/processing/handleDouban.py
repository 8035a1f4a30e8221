"""Import-compatibility stub for /root/reference/processing/handleDouban.py (see handleMovielens.py)."""


def prepareDouban(dataset_path_dict: dict, save_path: str):
    raise NotImplementedError(
        "prepareDouban is out of scope of the B200 hot-path drop-in: produce the pre-split CSVs with the "
        "reference's processing/ first")
