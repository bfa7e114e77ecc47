"""CPU pin of the exact parallel float32 prefix-sum algorithm behind the k-means++ pick kernel: the Python model
(tools/prefix_proto.py, the same summaries / composition / crossing rule as pp_pick_parallel_kernel) must reproduce the
sequential float32 running sum of pq.go:299-303 bit for bit, every 32-element checkpoint included."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import prefix_proto as P


def test_special_cases():
    assert P.special_cases() == []


def test_fuzz():
    assert P.fuzz(60) == 0
