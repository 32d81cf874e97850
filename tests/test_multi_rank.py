"""The N > 1 path on the CPU: two processes over gloo.  Inputs share no state, so ranks own disjoint
inputs and the data path has no collective; what is distributed is the partition, the barrier and the max-over-ranks time.
The per-rank "engine" here is the CPU oracle (a GPU is not available to this suite): the property checked is the one the
GPU path relies on - a rank's results do not depend on which other inputs run beside it."""
import os
import socket

import numpy as np
import pytest

from boondock_airband_b200 import abi, configs, sharding, synth


def test_partitions():
    assert sharding.inputs_of_rank(64, 8, 3) == list(range(3, 64, 8))  # BASELINE cfg 3: 8 inputs per GPU on 8 GPUs
    shards = [sharding.inputs_of_rank(13, 4, r) for r in range(4)]
    assert sharding.gather_shards(shards) == list(range(13)) and [len(s) for s in shards] == [4, 3, 3, 3]
    with pytest.raises(ValueError):
        sharding.gather_shards([[0, 1], [1, 2]])
    with pytest.raises(ValueError):
        sharding.inputs_of_rank(4, 2, 2)
    assert sharding.first_input_of_rank(512, 3) == 1536
    assert sharding.aggregate_msps(8, 10, 1_310_720_000, 0.06) == pytest.approx(8 * 10 * 1_310_720_000 / 0.06 / 1e6)
    assert sharding.max_over_ranks(1.5) == 1.5  # no process group: the rank's own time


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_inputs, seconds, out_dir):
    import torch
    import torch.distributed as dist
    from oracle.ba_oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.inputs_of_rank(n_inputs, world, rank)
        full = configs.cfg3(n_inputs)
        cfg = abi.EngineCfg(fft_size=full.fft_size, wave_rate=full.wave_rate, devices=[full.devices[i] for i in mine])
        o = Oracle(cfg)
        dist.barrier()
        for k, i in enumerate(mine):
            o.feed(k, synth.synth(full.devices[i], seconds, i, gate_on=0.2, gate_off=0.08))
        fake_elapsed = 1.0 + rank  # rank 1 is the slow one
        slowest = sharding.max_over_ranks(fake_elapsed, dist, torch.device("cpu"))
        assert slowest == float(world)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert sharding.gather_shards(gathered) == list(range(n_inputs))
        for k, i in enumerate(mine):
            np.save(os.path.join(out_dir, "in%d.npy" % i), np.stack([o.waveout(k, c) for c in range(len(cfg.devices[k].channels))]))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo_match_one_process(tmp_path, oracle_built):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    from oracle.ba_oracle import Oracle
    n_inputs, seconds, world = 3, 0.4, 2
    mp.spawn(_worker, args=(world, _free_port(), n_inputs, seconds, str(tmp_path)), nprocs=world, join=True)
    full = configs.cfg3(n_inputs)
    o = Oracle(full)  # one process, all inputs side by side
    for i in range(n_inputs):
        o.feed(i, synth.synth(full.devices[i], seconds, i, gate_on=0.2, gate_off=0.08))
    for i in range(n_inputs):
        got = np.load(tmp_path / ("in%d.npy" % i))
        want = np.stack([o.waveout(i, c) for c in range(16)])
        assert got.shape == want.shape and got.shape[1] >= 2 * full.wave_batch
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), "input %d depends on its neighbours" % i
