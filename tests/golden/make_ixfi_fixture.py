#!/usr/bin/env python
"""Hand-assembled faiss ``IndexFlatIP`` file: every byte is written by an explicit
``struct.pack`` below, field by field, following faiss' serializer — NOT by this repo's ``write_flat_ip_file`` — so that
``sdag_b200.retrieval.parse_flat_ip_file`` is checked against an independent statement of the layout.

faiss is not installable here (no network), so the layout is restated from its sources:
  impl/index_write.cpp  write_index():  ``uint32_t h = fourcc("IxFI"); WRITE1(h); write_index_header(idx, f);
                                          WRITEXBVECTOR(idxf->codes);``   (IndexFlat with METRIC_INNER_PRODUCT)
  impl/index_write.cpp  write_index_header(): ``WRITE1(idx->d); WRITE1(idx->ntotal); idx_t dummy = 1 << 20;
                                          WRITE1(dummy); WRITE1(dummy); WRITE1(idx->is_trained); WRITE1(idx->metric_type);
                                          if (idx->metric_type > 1) WRITE1(idx->metric_arg);``
  impl/io_macros.h      WRITEXBVECTOR(vec): ``size_t size = vec.size() / 4; WRITEANDCHECK(&size, 1);
                                          WRITEANDCHECK(vec.data(), size * 4);``  — ``codes`` is the byte vector holding the
                                          float32 rows, so ``size`` = number of float32 values = ntotal * d.
  fourcc("IxFI") = 'I' | 'x' << 8 | 'F' << 16 | 'I' << 24: the four characters in file order.
Types: int = int32, idx_t = int64, bool = 1 byte, MetricType = int32 enum (METRIC_INNER_PRODUCT = 0), size_t = uint64;
little-endian, written back to back without padding.

    python tests/golden/make_ixfi_fixture.py   ->  tests/golden/faiss_flat_ip_3x64.index (+ .docs_meta.jsonl)
"""
import json
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
D, NTOTAL = 64, 3


def vectors() -> np.ndarray:
    """Row r, column c = (r + 1) * 0.5 - c / 64: closed form, so the test can recompute the payload."""
    r = np.arange(NTOTAL, dtype=np.float32)[:, None]
    c = np.arange(D, dtype=np.float32)[None, :]
    return ((r + 1) * np.float32(0.5) - c / np.float32(64)).astype("<f4")


def main():
    x = vectors()
    blob = b"".join([
        b"IxFI",                              # 0   fourcc
        struct.pack("<i", D),                 # 4   d
        struct.pack("<q", NTOTAL),            # 8   ntotal
        struct.pack("<q", 1 << 20),           # 16  dummy
        struct.pack("<q", 1 << 20),           # 24  dummy
        struct.pack("<B", 1),                 # 32  is_trained
        struct.pack("<i", 0),                 # 33  metric_type = METRIC_INNER_PRODUCT
        struct.pack("<Q", NTOTAL * D),        # 37  size (float32 count)
        x.tobytes(),                          # 45  payload
    ])
    assert len(blob) == 45 + 4 * NTOTAL * D
    path = os.path.join(HERE, "faiss_flat_ip_3x64.index")
    with open(path, "wb") as f:
        f.write(blob)
    with open(os.path.join(HERE, "faiss_flat_ip_3x64.docs_meta.jsonl"), "w", encoding="utf-8") as f:
        for i in range(NTOTAL):  # the metadata format of dense.py:38-41, 119-120: one {"id", "text"} object per line
            f.write(json.dumps({"id": f"doc-{i}", "text": f"text of document {i}"}) + "\n")
    print(f"wrote {path} ({len(blob)} bytes)")


if __name__ == "__main__":
    main()
