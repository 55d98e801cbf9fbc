"""Frame sharding across ranks (one process per GPU).  Frames are independent units, exactly as the reference shards
them with DistributedSampler per rank (det3d/datasets/loader/build_loader.py:27-37; tools/dist_test.py:98-130):
frame f -> rank f mod world.  No data-path collective; results are merged with ONE fixed-size all_gather
(the reference pickles variable-size python objects, det3d/utils/dist/dist_common.py:48-88)."""
import torch
import torch.distributed as dist


def frames_for_rank(num_frames, rank, world):
    return list(range(rank, num_frames, world))


def gather_detections(local, num_frames, post_max, rank, world, device="cpu"):
    """local: {frame_index: (boxes [k,7], scores [k])}.  Returns the same mapping for all frames on every rank."""
    per_rank = (num_frames + world - 1) // world
    block = torch.zeros((per_rank, post_max, 8), dtype=torch.float32, device=device)
    meta = torch.full((per_rank, 2), -1, dtype=torch.int32, device=device)
    for slot, f in enumerate(frames_for_rank(num_frames, rank, world)):
        boxes, scores = local[f]
        k = boxes.shape[0]
        block[slot, :k, :7] = torch.as_tensor(boxes)
        block[slot, :k, 7] = torch.as_tensor(scores)
        meta[slot, 0], meta[slot, 1] = f, k
    if world > 1:
        blocks = [torch.zeros_like(block) for _ in range(world)]
        metas = [torch.zeros_like(meta) for _ in range(world)]
        dist.all_gather(blocks, block)
        dist.all_gather(metas, meta)
    else:
        blocks, metas = [block], [meta]
    out = {}
    for b, m in zip(blocks, metas):
        for slot in range(per_rank):
            f, k = int(m[slot, 0]), int(m[slot, 1])
            if f >= 0:
                out[f] = (b[slot, :k, :7].cpu().numpy(), b[slot, :k, 7].cpu().numpy())
    return out
