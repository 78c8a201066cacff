import os
import sys

from helpers import ROOT


def test_brief_tables_are_the_generated_ones():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_brief_pairs
    txt = gen_brief_pairs.render()
    for rel in ("oracle/brief_pairs.inc", "lvt_b200/csrc/brief_pairs.inc"):
        assert open(os.path.join(ROOT, rel)).read() == txt
    p = gen_brief_pairs.pairs()
    assert len(p) == 256 and all(-24 <= v <= 24 for q in p for v in q)
