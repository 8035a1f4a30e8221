"""Import-compatibility stub for /root/reference/processing/handleMovielens.py.

main.py:13 imports prepareMovieLens unconditionally but only calls it when the pre-split CSVs are
missing (main.py:41-46).  Dataset ETL (quantile user filter, LabelEncoder, 80/10/10 split,
Word2Vec side features) is offline Python with no data-parallel math and is out of scope for
the B200 hot-path build; it also needs gensim/jieba/nltk.  Run the reference's
own processing/ once to produce the CSVs, then point const.py at them."""


def prepareMovieLens(dataset_path_dict: dict, save_path: str):
    raise NotImplementedError(
        "prepareMovieLens is out of scope of the B200 hot-path drop-in: produce filter_rating.csv / "
        "train_data.csv / val_data.csv / test_data.csv / *_features.csv with the reference's processing/ first")
