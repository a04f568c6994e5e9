"""N > 1 host logic on CPU: world_size-2 gloo run of the sharding helpers used by bench.py.
Images are independent, so the union of the ranks' results must equal the single-rank result byte for byte."""
import os, sys, subprocess, json, tempfile
import pytest
import harness as H

WORKER = r'''
import os, sys, json, hashlib
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests")); sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import numpy as np, torch.distributed as dist
import harness as H, cpu_encode as CE
from cavif_rs_b200.shard import shard_indices, broadcast_tables, gather_stats
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
tbl = H.rate_tables(121) if rank == 0 else np.zeros(H.RATE_TABLES_BYTES // 2, np.uint16)
tbl = broadcast_tables(tbl, rank)
mine = shard_indices(5, rank, world)
digests = {}
for i in mine:
    img = H.synth_image(72, 48, i)
    r = CE.encode_rgba(img, 80.0, 6, 8)
    digests[i] = hashlib.sha1(r["color_obu"]).hexdigest()
stats = gather_stats([float(len(mine)), float(sum(len(d) for d in digests.values()))])
json.dump({"rank": rank, "tbl": hashlib.sha1(tbl.tobytes()).hexdigest(), "digests": digests, "stats": stats},
          open(os.path.join(%(out)r, "r%%d.json" %% rank), "w"))
dist.destroy_process_group()
'''


def test_two_ranks_cover_the_batch_and_match_single_rank():
    import hashlib, numpy as np
    sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
    import cpu_encode as CE
    H.oracle(); H.hostlib()
    with tempfile.TemporaryDirectory() as td:
        script = os.path.join(td, "w.py")
        open(script, "w").write(WORKER % {"root": H.ROOT, "out": td})
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", "29731", script], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        res = [json.load(open(os.path.join(td, "r%d.json" % k))) for k in range(2)]
    assert res[0]["tbl"] == res[1]["tbl"] == hashlib.sha1(H.rate_tables(121).tobytes()).hexdigest()
    merged = {}
    for x in res:
        merged.update({int(k): v for k, v in x["digests"].items()})
    assert sorted(merged) == [0, 1, 2, 3, 4]
    assert set(res[0]["digests"]) == {"0", "2", "4"} and set(res[1]["digests"]) == {"1", "3"}
    for i in range(5):
        ref = CE.encode_rgba(H.synth_image(72, 48, i), 80.0, 6, 8)
        assert merged[i] == hashlib.sha1(ref["color_obu"]).hexdigest()
    assert res[0]["stats"][0][0] == 3.0 and res[0]["stats"][1][0] == 2.0
