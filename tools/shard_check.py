"""torchrun entry: subject-sharded preprocessing + NCCL stream exchange on WORLD_SIZE GPUs must give every rank,
bit for bit, what single-process preprocessing gives.  Used by tests/test_gpu_parallel.py (needs >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/shard_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    from multimodalsignal_b200 import preprocess as pp, synth
    from oracle import preprocess_oracle as po
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sids = ["S2", "S3", "S4", "S5", "S6"]
    subs = [synth.make_subject(s, i, minutes=synth.SHORT_MINUTES, protocol=synth.SHORT_PROTOCOL) for i, s in enumerate(sids)]
    items = [(s.sid, s.as_pickle_dict, po.apply_subject_quirk(s.sid, s.protocol)) for s in subs]
    got = pp.preprocess_subjects_sharded(items, 64, include_wrist=True)
    assert list(got) == sids
    for s, (sid, data, proto) in zip(subs, items):
        ref = pp.preprocess_subject(sid, data(), proto, 64, include_wrist=True)
        g = got[sid]
        assert torch.equal(g.streams, ref.streams), sid
        assert np.array_equal(g.starts_host, ref.starts_host) and np.array_equal(g.labels, ref.labels)
        assert torch.equal(g.starts, ref.starts) and g.window == ref.window and g.channel_names == ref.channel_names
    dist.barrier()
    if rank == 0:
        print("shard_check ok")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
