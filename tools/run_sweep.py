"""BASELINE.json configs[4]: all six emulators over N parameter sets (default 100 M), streamed in 1 M-row chunks.
Single process = one GPU; under torchrun each rank takes a contiguous range of chunks (no collective)."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402
import cosmopower_b200 as cp  # noqa: E402
from cosmopower_b200 import standins, sweep  # noqa: E402


def main():
    n_total = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 100_000_000
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    models = {}
    for name in standins.ALL_MODELS:
        kind, path = standins.model_path(name)
        models[name] = getattr(cp, kind)(restore=True, restore_filename=path, device=local)
    res = sweep.sweep(models, n_total, chunk_rows=1 << 20, seed=1, device=local, ten_to=False, checksum=("--no-checksum" not in sys.argv),
                      world_size=world, rank=rank)
    for name, r in res.items():
        r["spectra_per_s"] = r["rows"] / r["seconds"]
    print(json.dumps({"rank": rank, "world": world, "n_total": n_total, "results": res}), flush=True)


if __name__ == "__main__":
    main()
