"""Two ranks on two GPUs: the data-parallel entry point against the same work done on one GPU.

Every rank encodes its shard of a RAGGED batch (5 / 3 / 4 / 5 tiles) and the projected tokens travel to rank 0
by both transports (peer stores into rank 0's symmetric buffer, and the NCCL all-gather).  Rank 0 also encodes
each rank's shard by itself, in the shard's own shapes (same kernel forms, deterministic kernels), and splices
them: the sharded result must equal that BIT FOR BIT, on both transports; against the plain single-process call
(other GEMM tilings at 17 tiles) it must agree within the bf16 tolerance.  Skipped with fewer than 2 GPUs."""
import os

import numpy as np
import pytest
import torch

from helpers import PINPOINTS_C3, cos_rows, synth_image

pytestmark = pytest.mark.gpu

SIZES = [(1000, 900), (637, 336), (336, 900), (700, 650)]          # -> 5, 3, 4, 5 tiles under PINPOINTS_C3


def _batch():
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(3, 32000, (4, 48), generator=g)
    for b, pos in enumerate([3, 40, 0, 17]):
        ids[b, pos] = -200
    ids[1, 44:] = 2
    ids[3, 30:] = 2
    mask = (ids != 2).long()
    labels = ids.clone()
    labels[:, :8] = -100
    return ids, mask, labels


def _worker(rank, world, port, lut, q):
    try:
        import torch.distributed as dist
        import vision_zephyr_b200 as vz
        from vision_zephyr_b200.dist import shard_images
        from vision_zephyr_b200.runtime import VisionEmbeddingPath, random_init_
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        path = random_init_(VisionEmbeddingPath(device=dev), seed=0)          # same seed -> same weights on every rank
        imgs = [torch.from_numpy(synth_image(60 + i, w, h)).to(dev) for i, (w, h) in enumerate(SIZES)]
        pb_all = vz.process_any_resolution_images(imgs, PINPOINTS_C3, lut, out_mode="patches")
        tiles = pb_all.tiles_per_image
        assert tiles == [5, 3, 4, 5], tiles
        bounds = shard_images(tiles, world)
        lo, hi = bounds[rank]
        pb = vz.process_any_resolution_images(imgs[lo:hi], PINPOINTS_C3, lut, out_mode="patches")
        ids, mask, labels = (t.to(dev) for t in _batch())
        res = {}
        for env, name in (("0", "nccl_all_gather"), ("1", "peer_store")):
            os.environ["VZ_PEER_GATHER"] = env
            out = path.prepare_inputs_labels_for_multimodal_sharded(ids, None, mask, None, labels, pb, tiles, SIZES)
            assert path.last_transport == name, (path.last_transport, name)
            torch.cuda.synchronize()
            if rank == 0:
                res[name] = [t.clone() for t in (out[4], out[5], out[2])]
            else:
                assert out[4] is None
        ok = {}
        if rank == 0:
            # every shard by itself on this GPU, in the shard's own shapes, then one splice
            ctx = path._plan_splice(ids, mask, labels, tiles, SIZES)
            tower, proj = path.get_vision_tower(), path.get_model().mm_projector
            parts, t0 = [], 0
            for (a, b) in bounds:
                n = sum(tiles[a:b])
                feats = tower.encode_patches(pb_all.patches[t0 * 576:(t0 + n) * 576], pre_norm=proj.pre_norm_params())
                parts.append(path._project_shard(ctx, feats, tiles, a, b).clone())
                t0 += n
            ref = path._splice(ctx, torch.cat(parts), None, mask, None, labels)
            torch.cuda.synchronize()
            for name, (e, l, m) in res.items():
                ok[name] = bool(torch.equal(e, ref[4]) and torch.equal(l, ref[5]) and torch.equal(m, ref[2]))
            ok["transports_equal"] = bool(torch.equal(res["peer_store"][0], res["nccl_all_gather"][0]))
            plain = path.prepare_inputs_labels_for_multimodal(ids, None, mask, None, labels, pb_all, SIZES)
            a, b = res["peer_store"][0].float().cpu().numpy(), plain[4].float().cpu().numpy()
            ok["plain_shape"] = a.shape == b.shape
            vis_rows = (plain[5] == -100).cpu().numpy() & (plain[2] != 0).cpu().numpy()
            ok["plain_cos"] = float(cos_rows(a[vis_rows], b[vis_rows]).min())
            ok["plain_err"] = float(np.abs(a - b).max())
            ok["ints_equal"] = bool(torch.equal(res["peer_store"][1], plain[5]) and torch.equal(res["peer_store"][2], plain[2]))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, ok, None))
    except Exception as e:  # surface the failure in the parent instead of a timeout
        import traceback
        q.put((rank, None, traceback.format_exc()))


def test_two_rank_sharded_equals_unsharded_on_both_transports(golden_dir):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    lut = np.load(f"{golden_dir}/golden_pixels.npz")["lut"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lut, q)) for r in range(2)]
    [p.start() for p in procs]
    got = sorted((q.get(timeout=600) for _ in range(2)), key=lambda x: x[0])
    [p.join(60) for p in procs]
    for rank, ok, err in got:
        assert err is None, f"rank {rank}:\n{err}"
    ok = got[0][1]
    print(ok)
    assert ok["nccl_all_gather"] and ok["peer_store"] and ok["transports_equal"]
    assert ok["plain_shape"] and ok["ints_equal"] and ok["plain_cos"] >= 0.9995 and ok["plain_err"] <= 0.1
