"""__graft_entry__._compile: a built file is fresh against the CONTENT of its sources and the compiler flags (modification times
and the location of the tree do not matter), and concurrent callers (the ranks of a torchrun launch) never see a partial file."""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def _cmd(src, log):
    # the "compiler": appends a line to `log`, takes a moment, copies the source
    return lambda out: [sys.executable, "-c",
                        "import shutil, sys, time; open(sys.argv[3], 'a').write('x\\n'); time.sleep(0.3); shutil.copy(sys.argv[1], sys.argv[2])",
                        src, out, log]


def _runs(log):
    return len(open(log).read().split()) if os.path.exists(log) else 0


def test_fresh_by_content_not_by_time(tmp_path):
    src, tgt, log = str(tmp_path / "a.cu"), str(tmp_path / "a.out"), str(tmp_path / "log")
    open(src, "w").write("one")
    ge._compile(tgt, [src], _cmd(src, log))
    assert open(tgt).read() == "one" and _runs(log) == 1
    os.utime(src, (time.time() + 1000, time.time() + 1000))        # a newer modification time alone does not rebuild
    ge._compile(tgt, [src], _cmd(src, log))
    assert _runs(log) == 1
    open(src, "w").write("two")
    ge._compile(tgt, [src], _cmd(src, log))
    assert open(tgt).read() == "two" and _runs(log) == 2
    os.remove(tgt)                                                  # a missing target is rebuilt whatever the stamp says
    ge._compile(tgt, [src], _cmd(src, log))
    assert open(tgt).read() == "two" and _runs(log) == 3
    ge._compile(tgt, [src], _cmd(src, log), force=True)
    assert _runs(log) == 4
    assert not [f for f in os.listdir(str(tmp_path)) if ".tmp" in f]


def _worker(args):
    src, tgt, log = args
    ge._compile(tgt, [src], _cmd(src, log))
    return open(tgt).read()


def test_concurrent_callers_build_once(tmp_path):
    src, tgt, log = str(tmp_path / "b.cu"), str(tmp_path / "b.out"), str(tmp_path / "log")
    open(src, "w").write("payload" * 1000)
    with mp.get_context("spawn").Pool(4) as pool:
        outs = pool.map(_worker, [(src, tgt, log)] * 4)
    assert all(o == "payload" * 1000 for o in outs) and _runs(log) == 1


def test_digest_ignores_where_the_tree_sits():
    a = ge._digest([], ["nvcc", "-o", "@", os.path.join(ge.ROOT, "nspeech_b200", "csrc", "nspeech_b200.cu")])
    b = ge._digest([], ["nvcc", "-o", "@", os.path.join(".", "nspeech_b200", "csrc", "nspeech_b200.cu")])
    assert a == b
