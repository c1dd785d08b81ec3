"""Save / load bandwidth of the sharded on-disk index format: a 20 000-page ColPali shard (5.3 GB) is written
as 4 shard directories through `lis_index_save_rows` and read back through `lis_index_load_rows` (pinned double buffer, parallel
pread/pwrite), on whatever storage the box gives under the target directory; plus the world-2-style load (two halves)."""
import importlib, json, shutil, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
lis = importlib.import_module("multi-modal_colpali_b200")
pages = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000
for target in ("/dev/shm/lis_bw", "/tmp/lis_bw"):
    d = Path(target)
    try:
        shutil.rmtree(d, ignore_errors=True)
        idx = lis.LateInteractionIndex(pages * 1030, pages)
        idx.fill_synthetic(pages, 1030, seed=3)
        q = torch.nn.functional.normalize(torch.randn(1, 16, 128, generator=torch.Generator().manual_seed(1)), dim=-1).to(torch.bfloat16)
        want = idx.search(q, 10)
        gb = idx.num_rows * 256 / 1e9
        t0 = time.perf_counter(); idx.save(d, shards=4); t_save = time.perf_counter() - t0
        idx.close()
        rec = {"dir": target, "pages": pages, "GB": round(gb, 2), "save_s": round(t_save, 3), "save_GBps": round(gb / t_save, 2)}
        for label in ("load_1", "load_2"):       # second load: page cache warm
            t0 = time.perf_counter(); back = lis.LateInteractionIndex.load(d); torch.cuda.synchronize(); t = time.perf_counter() - t0
            got = back.search(q, 10)
            assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
            back.close()
            rec[label + "_s"] = round(t, 3); rec[label + "_GBps"] = round(gb / t, 2)
        t0 = time.perf_counter()
        halves = [lis.LateInteractionIndex.load(d, shard_ids=list(range(a, b))) for a, b in lis.assign_shards([1, 1, 1, 1], 2)]
        torch.cuda.synchronize(); rec["load_two_halves_s"] = round(time.perf_counter() - t0, 3)
        [h.close() for h in halves]
        print(json.dumps(rec), flush=True)
    except Exception as exc:   # a target directory may not exist / be too small on a given box
        print(json.dumps({"dir": target, "error": repr(exc)[:200]}), flush=True)
    finally:
        shutil.rmtree(d, ignore_errors=True)
