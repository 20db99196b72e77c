import torch

from tf_yarn_b200 import data


def test_batch_shuffle_repeat_take():
    x = torch.arange(10).float().unsqueeze(1)
    y = torch.arange(10)
    ds = data.Dataset.from_tensor_slices((x, y)).shuffle(5, seed=0).batch(4).repeat(2)
    batches = list(ds)
    assert [b[0].shape[0] for b in batches] == [4, 4, 2, 4, 4, 2]
    assert sorted(torch.cat([b[1] for b in batches[:3]]).tolist()) == list(range(10))
    assert len(list(ds.take(2))) == 2


def test_filter_map_dict_features():
    feats = {"a": torch.arange(6).float(), "b": torch.ones(6)}
    ds = data.Dataset.from_tensor_slices((feats, torch.arange(6)))
    ds = ds.filter(lambda f, y: y % 2 == 0).map(lambda f, y: ({"a": f["a"] * 2, "b": f["b"]}, y)).batch(2)
    out = list(ds)
    assert out[0][0]["a"].tolist() == [0.0, 4.0] and out[0][1].tolist() == [0, 2]
    assert len(out) == 2


def test_csv_dataset(tmp_path):
    p = tmp_path / "d.csv"
    p.write_text("h1;h2;h3\n1.5;2;a\n2.5;3;b\n")
    rows = list(data.CsvDataset(str(p), [0.0, 0, ""], header=True, field_delim=";"))
    assert rows == [(1.5, 2, "a"), (2.5, 3, "b")]


def test_shard_and_repeat_forever():
    ds = data.Dataset.range(10).shard(2, 1)
    assert list(ds) == [1, 3, 5, 7, 9]
    it = iter(data.Dataset.range(3).repeat())
    assert [next(it) for _ in range(7)] == [0, 1, 2, 0, 1, 2, 0]
