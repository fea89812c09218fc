#!/usr/bin/env python3
"""Regenerate tests/golden/colex_checksums.json: per-chunk checksums of the rank ids of ALL C(52,7) hands in
colexicographic order, computed with the CPU oracle (whose evaluator is pinned against the reference by eval_tables.npz /
eval_cases.json).  ~1 s on 8 threads.   python tests/golden/make_colex_checksums.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle  # noqa: E402

CHUNK = 1 << 21
cs = oracle.colex_checksums(CHUNK, threads=os.cpu_count() or 1)
tot = cs.sum(0)
assert tot[2:].tolist() == [23294460, 58627800, 31433400, 6461620, 6180020, 4047644, 3473184, 224848, 41584]
json.dump({"source": "oracle.colex_checksums(1<<21): oracle_calc_score-derived rank ids over all C(52,7) hands in colex order",
           "chunk": CHUNK, "n_hands": oracle.N_HANDS_7,
           "columns": ["sum_rank_id", "sum_rank_id_times_((index%65521)+1)"] + oracle.CAT_NAMES,
           "total": tot.tolist(), "chunks": cs.tolist()},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "colex_checksums.json"), "w"))
print("written; sum of all rank ids =", int(tot[0]))
