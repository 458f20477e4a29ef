"""Generates the small storage fixtures under tests/golden/ (run once, outputs committed):
  vectors_parquetgo.parquet  the `vectors.parquet` of a persisted collection in the physical layout
                             xitongsys/parquet-go derives from ParquetVectorRecord's struct tags
                             (reference pkg/persistence/parquet.go:16-20): REQUIRED fields, three-level
                             LIST (`list` / `element`), dictionary on `id`, SNAPPY, data page v1
  index_arrow_hnsw.arrow     the Arrow IPC file ArrowHNSWIndex.Save writes (reference index/arrow_hnsw.go:153-197):
                             `id: utf8`, `vector: fixed_size_list<float32>[dim]`, several record batches
The reference's own writers are Go and cannot run in this image; the layouts are reproduced with pyarrow and
pinned by tests/test_ingest.py (schema text, encodings, row rules, and an oracle search over the loaded rows)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from quiver_b200 import ingest  # noqa: E402

rng = np.random.default_rng(20260101)
n, d = 300, 12
ids = [f"vec-{i:04d}" for i in range(n)]
vec = rng.standard_normal((n, d)).astype(np.float32)
md = [{"category": f"cat{i % 4}", "lang": "en" if i % 3 else "de"} for i in range(n)]
# rows the reader must drop (parquet.go:133-166): an empty id, an empty vector; and one unparsable document
ids_p = ids[:100] + [""] + ids[100:200] + ["empty-vector"] + ids[200:]
vec_p = list(vec[:100]) + [vec[0]] + list(vec[100:200]) + [[]] + list(vec[200:])
md_p = md[:100] + [{"category": "ghost"}] + md[100:200] + [{"category": "ghost"}] + md[200:]
md_p[7] = "not json"
ingest.save_parquet(os.path.join(HERE, "vectors_parquetgo.parquet"), ids_p, vec_p, md_p, parquet_go_layout=True)
ingest.save_arrow_ipc(os.path.join(HERE, "index_arrow_hnsw.arrow"), ids, vec, batch_rows=128)
np.save(os.path.join(HERE, "storage_fixture_vectors.npy"), vec)
json.dump({"ids": ids, "metadata": md, "dropped": ["", "empty-vector"], "unparsable_metadata_row": 7},
          open(os.path.join(HERE, "storage_fixture_rows.json"), "w"))
print("fixtures written")
