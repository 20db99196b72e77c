import pyarrow as pa
import pyarrow.parquet as pq
import pytest

from tf_yarn_b200.pytorch.parquet_dataset import ParquetDataset


@pytest.fixture
def parquet_dir(tmp_path):
    for f in range(2):
        table = pa.table({"x": list(range(f * 100, f * 100 + 100)), "y": [float(i) for i in range(100)]})
        pq.write_table(table, tmp_path / f"part-{f}.parquet")
    (tmp_path / "_SUCCESS").write_text("")
    return str(tmp_path)


def test_num_samples_and_len(parquet_dir):
    ds = ParquetDataset(parquet_dir, batch_size=10, rank=0, world_size=2)
    assert ds.num_samples == 200
    assert len(ds) == 200 // 10 // 2


def test_each_rank_gets_a_disjoint_contiguous_slice(parquet_dir):
    seen = []
    for rank in range(2):
        ds = ParquetDataset(parquet_dir, batch_size=10, columns=["x"], rank=rank, world_size=2)
        rows = [v for batch in ds for v in batch.column("x").to_pylist()]
        seen.append(rows)
        # per file: 10 batches, last dropped -> 9 -> 4 per rank
        assert len(rows) == 2 * 4 * 10
    assert not set(seen[0]) & set(seen[1])
    assert seen[0][:10] == list(range(10)) and seen[1][:10] == list(range(40, 50))
