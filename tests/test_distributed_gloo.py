"""CPU, world_size 2 over gloo: the data-parallel plumbing (sharding, result gather, gradient / normaliser all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port() -> int:
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def _worker(rank: int, world: int, port: int, results) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from allophant_b200.dataset_processing import Batch
        from allophant_b200.distributed import allreduce_gradients, gather_by_index, global_label_count, shard_batch

        lengths = torch.tensor([100, 900, 500, 300, 700])
        audio = torch.arange(5 * 900, dtype=torch.float32).view(5, 900)
        shard, indices = shard_batch(Batch(audio, lengths, torch.arange(5)), rank, world)
        assert shard.audio_features.shape[1] == int(shard.lengths.max())
        local = {index: float(shard.audio_features[i, 0]) for i, index in enumerate(indices)}
        merged = gather_by_index(local)
        assert sorted(merged) == [0, 1, 2, 3, 4]
        assert all(merged[i] == float(audio[i, 0]) for i in range(5))
        # gradients: sum over ranks, bucketed
        params = [torch.nn.Parameter(torch.zeros(1000)), torch.nn.Parameter(torch.zeros(3, 7))]
        for p in params:
            p.grad = torch.full_like(p, float(rank + 1))
        issued = allreduce_gradients(params, bucket_bytes=2048)
        assert issued == 2 and all(torch.equal(p.grad, torch.full_like(p, 3.0)) for p in params)
        total = global_label_count([torch.tensor([3, 4]) + rank, torch.tensor([1])])
        assert float(total) == (8 + 10)
        # overlapped reducer: flat groups are reduced in place, small tensors are packed into one bucket
        from allophant_b200.distributed import GradientReducer

        reducer = GradientReducer()
        flat = torch.full((64,), float(rank + 1))
        views = {"a": flat[:16].view(4, 4), "b": flat[16:]}
        reducer.submit(flat, views)
        packed = reducer.submit_tensors({"w": torch.full((3, 5), float(rank)), "b": torch.full((7,), 2.0 * rank)})
        reducer.finish()
        assert reducer.issued == 2 and not reducer.works
        assert torch.equal(views["a"], torch.full((4, 4), 3.0)) and torch.equal(views["b"], torch.full((48,), 3.0))
        assert packed["w"].shape == (3, 5) and torch.equal(packed["w"], torch.full((3, 5), 1.0))
        assert torch.equal(packed["b"], torch.full((7,), 2.0))
        # bf16 on the wire: half the bytes, the fp32 buffer (and its views) receive the rounded sum
        compressed = GradientReducer(wire_dtype=torch.bfloat16)
        flat = torch.full((32,), 1.0 + rank)
        view = flat[:8]
        compressed.submit(flat, {"v": view})
        compressed.finish()
        assert compressed.bytes == 64 and torch.equal(view, torch.full((8,), 3.0)) and flat.dtype == torch.float32
        results[rank] = indices
    finally:
        dist.destroy_process_group()


def test_sharding_gather_and_allreduce_world2():
    manager = mp.Manager()
    results = manager.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert sorted(results[0] + results[1]) == [0, 1, 2, 3, 4]
    # longest-first serpentine: 900,700 | 700->rank1 ... balanced sample counts
    from allophant_b200.distributed import shard_indices

    shards = shard_indices([100, 900, 500, 300, 700], 2)
    loads = [sum([100, 900, 500, 300, 700][i] for i in shard) for shard in shards]
    assert abs(loads[0] - loads[1]) <= 300
