#!/usr/bin/env python
"""Export / reload of the production on-disk format at BASELINE config-3 size (60 M interactions, 1.72 M items): wall
times of vmis_index_to_avro and vmis_index_from_avro (the reference notes 161 s single-threaded for the item index
alone, vmis_index.rs:201) and an identity check of the reloaded index.  GPU box: python tools/avro_scale.py [dir]"""
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import serenade_b200 as sb  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/vmis_avro_scale"
shutil.rmtree(out, ignore_errors=True)
t0 = time.time()
items, off, ts = sb.synth_sessions(42, 1_760_000, 11_556_000)
gix = sb.VMISIndex.from_sessions(items, off, ts, 1502, 34, 2.0, device=0)
t1 = time.time()
gix.to_avro(out, "deflate", 16)
t2 = time.time()
size = sum(os.path.getsize(os.path.join(d, f)) for d, _, fs in os.walk(out) for f in fs)
back = sb.VMISIndex.new(out, device=0)
t3 = time.time()
q = sb.synth_queries(43, 1_760_000, 1 << 16, 4)
a = sb.predict_batch(gix, q, 288, 1502, 21)
b = sb.predict_batch(back, q, 288, 1502, 21)
same = all(np.array_equal(x, y) for x, y in zip(a, b))
print(f"synth+build {t1 - t0:.1f}s  export {t2 - t1:.1f}s ({size / 1e6:.0f} MB deflate, 16+16 files)  "
      f"load+check+upload {t3 - t2:.1f}s  info {back.prebuilt_info()}  identical_predictions {same}")
shutil.rmtree(out, ignore_errors=True)
