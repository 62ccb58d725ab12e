"""Run under torchrun on G >= 2 GPUs of one box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [shape]
Checks that the row-partitioned propagation (NCCL all-gather per layer) and the item-sharded top-K
(NCCL all-gather of partial lists + merge) are BIT-IDENTICAL to the single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import gnn_recommendations_b200 as g  # noqa: E402
from gnn_recommendations_b200.dist import (PeerExchange, RowPartition, ShardedLightGCN, full_rank_topk_sharded,  # noqa: E402
                                           gather_rows, item_shard, lightgcn_propagate_fused,
                                           lightgcn_propagate_sharded)
from gnn_recommendations_b200.evaluator import ground_truth_dict, seen_csr  # noqa: E402
from gnn_recommendations_b200.synthetic import SHAPES, synth_split  # noqa: E402


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "C1"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    sp = synth_split(shape, 42)
    nu, ni, _, d, L = SHAPES[shape]
    full = g.NormAdjCSR.from_pairs(*sp["train"], nu, ni, device=dev)
    gen = torch.Generator().manual_seed(0)
    x0 = (torch.randn(nu + ni, d, generator=gen) * 0.1).to(dev)
    single = g.lightgcn_propagate(full, x0, L)

    part = RowPartition(full.indptr, world)
    local_csr = part.local_csr(full, rank)
    x0_mine = part.take_rows(x0, rank)
    mine = lightgcn_propagate_sharded(local_csr, part, rank, x0_mine, L)
    gathered = gather_rows(part, rank, mine)
    ok_prop = bool(torch.equal(gathered, single))
    # fused SpMM + all-gather over peer memory (NVLink P2P stores from the epilogue)
    ex = PeerExchange(part, d, dev)
    ok_fused = True
    for _ in range(3):                                   # repeated calls reuse the two buffers
        mine_f = lightgcn_propagate_fused(local_csr, ex, x0_mine, L)
        ok_fused = ok_fused and bool(torch.equal(mine_f, mine))

    gt = ground_truth_dict(sp["test"])
    eu = sorted(gt)
    ip, it = seen_csr(eu, nu, sp["train"], sp["valid"])
    want = g.full_rank_topk(single[:nu], single[nu:], eu, ip, it, 20)
    lo, hi = item_shard(ni, world, rank)
    got = full_rank_topk_sharded(gathered[:nu], gathered[nu + lo: nu + hi], lo, hi, eu, ip, it, 20, world, n_splits=2)
    ok_topk = bool(torch.equal(got, want))
    # row-partitioned training steps (ShardedLightGCN) vs the single-GPU step on the full graph
    import time

    from gnn_recommendations_b200.losses import bpr_fused
    from gnn_recommendations_b200.optim import fused_clip_adam_step
    B, n_steps = 512, 20
    sharded = ShardedLightGCN(local_csr, part, rank, x0_mine, nu, L, exchange=ex)
    w_full = torch.nn.Parameter(x0.clone())
    opt = torch.optim.Adam([w_full], lr=1e-3, weight_decay=1e-4)
    rng = np.random.default_rng(11)
    batches = [(torch.from_numpy(rng.integers(0, nu, B)).to(dev), torch.from_numpy(rng.integers(0, ni, B)).to(dev),
                torch.from_numpy(rng.integers(0, ni, B)).view(-1, 1).to(dev)) for _ in range(n_steps)]
    ok_train, t_sh, t_1 = True, 0.0, 0.0
    for users, pos, neg in batches:
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        loss_sh = sharded.train_step(users, pos, neg)
        torch.cuda.synchronize(); dist.barrier(); t_sh += time.perf_counter() - t0
        t0 = time.perf_counter()
        opt.zero_grad()
        xp = g.lightgcn._LightGCNPropagate.apply(w_full, full, L)
        loss_1 = bpr_fused(xp, nu, users, pos, neg)
        loss_1.backward()
        fused_clip_adam_step(opt, 1.0)
        torch.cuda.synchronize(); t_1 += time.perf_counter() - t0
        ok_train = ok_train and abs(loss_sh - float(loss_1)) <= 1e-6 * abs(float(loss_1))
    ok_train = ok_train and bool(torch.allclose(sharded.gathered_weight(), w_full.detach(), rtol=1e-4, atol=2e-6))
    # ---- user-owner (1.5-D) mode: item rows = partial sums reduced + broadcast over peer memory (1e-5 parity)
    from gnn_recommendations_b200.dist import (BipartitePartition, ItemExchange, UserOwnerLightGCN,
                                               lightgcn_propagate_user_owner)
    bp = BipartitePartition(nu, ni, world)
    a_u, a_i = bp.local_csrs(full, rank)
    # partitioned build: the same two matrices straight from this rank's pairs + one all-reduce of the item degrees
    from gnn_recommendations_b200.dist import build_user_owner_csrs
    b_u, b_i = build_user_owner_csrs(bp, rank, *sp["train"], device=dev, long_threshold=full.long_threshold,
                                     item_degree_allreduce=dist.all_reduce)
    ok_build = all(torch.equal(x.indptr, y.indptr) and torch.equal(x.indices, y.indices) and
                   torch.equal(x.vals.view(torch.int32), y.vals.view(torch.int32)) for x, y in ((b_u, a_u), (b_i, a_i)))
    a_u, a_i = b_u, b_i                                  # everything below runs on the partition-built matrices
    iex = ItemExchange(bp, d, dev)
    lo_b, hi_b = bp.item_range(rank)
    xi_blk = torch.zeros((bp.item_block, d), device=dev)
    xi_blk[: hi_b - lo_b] = x0[nu + lo_b: nu + hi_b]
    xu_mine = bp.take_users(x0[:nu], rank)
    ok_uo = True
    scale = float(single.abs().max())
    for _ in range(3):
        ou, oi = lightgcn_propagate_user_owner(a_u, a_i, iex, xu_mine, xi_blk, L)
        err = max(float((ou - single[:nu][rank::world]).abs().max()),
                  float((oi - single[nu + lo_b: nu + hi_b]).abs().max()) if hi_b > lo_b else 0.0)
        ok_uo = ok_uo and err <= 1e-5 * scale
    uo = UserOwnerLightGCN(a_u, a_i, iex, xu_mine, xi_blk, L)
    w2 = torch.nn.Parameter(x0.clone())
    opt2 = torch.optim.Adam([w2], lr=1e-3, weight_decay=1e-4)
    ok_uo_train, t_uo = True, 0.0
    for users, pos, neg in batches[:10]:
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        loss_uo = uo.train_step(users, pos, neg)
        torch.cuda.synchronize(); dist.barrier(); t_uo += time.perf_counter() - t0
        opt2.zero_grad()
        loss_1 = bpr_fused(g.lightgcn._LightGCNPropagate.apply(w2, full, L), nu, users, pos, neg)
        loss_1.backward()
        fused_clip_adam_step(opt2, 1.0)
        ok_uo_train = ok_uo_train and abs(loss_uo - float(loss_1)) <= 1e-6 * abs(float(loss_1))
    ok_uo_train = ok_uo_train and bool(torch.allclose(uo.users.detach(), w2.detach()[:nu][rank::world], rtol=1e-4, atol=2e-6))
    if hi_b > lo_b:
        ok_uo_train = ok_uo_train and bool(torch.allclose(uo.items.detach()[: hi_b - lo_b], w2.detach()[nu + lo_b: nu + hi_b],
                                                          rtol=1e-4, atol=2e-6))
    flags = torch.tensor([int(ok_prop), int(ok_topk), int(ok_fused), int(ok_train), int(ok_uo), int(ok_uo_train),
                          int(ok_build)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multigpu_check shape={shape} world={world} rows/rank={[part.n_local(r) for r in range(world)]} "
              f"propagation_bit_identical={bool(flags[0])} topk_bit_identical={bool(flags[1])} "
              f"fused_peer_exchange_bit_identical={bool(flags[2])} sharded_training_matches={bool(flags[3])} "
              f"user_owner_propagation_within_1e-5={bool(flags[4])} user_owner_training_matches={bool(flags[5])} "
              f"user_owner_partitioned_build_bit_identical={bool(flags[6])} "
              f"train_step_ms: sharded {1e3 * t_sh / n_steps:.3f} / user-owner {1e3 * t_uo / 10:.3f} vs 1 GPU "
              f"{1e3 * t_1 / n_steps:.3f}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if bool(flags.min()) else 1)


if __name__ == "__main__":
    main()
