"""Multi-GPU parity check, one process per GPU (torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29544 tests/multigpu_check.py

Every rank partitions the golden (reference-built) hierarchy, uploads ITS share and runs the
distributed operators / V-cycle / PCG over NCCL; rank 0 gathers the pieces and compares with the
CPU oracle's emulation of the same partition (same float halo truncation, same agglomeration).
Also run (on one box with >= 2 GPUs) by tests/test_multigpu.py.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.oracle import Oracle  # noqa: E402
from saena_b200 import native  # noqa: E402
from saena_b200.distributed import exchange_nccl_id, setup_p2p_halo  # noqa: E402
from saena_b200.hierarchy import KIND_A, KIND_P, KIND_R, partition_hierarchy  # noqa: E402
from tests.util import GOLDEN, TOL_HIST, TOL_OP, Golden, rel  # noqa: E402


def gather(x, sizes, rank, world):
    """concatenate per-rank numpy vectors on every rank"""
    out = [torch.zeros(s, dtype=torch.float64, device="cuda") for s in sizes]
    dist.all_gather(out, torch.from_numpy(np.ascontiguousarray(x)).cuda()) if len(set(sizes)) == 1 else None
    if len(set(sizes)) != 1:
        mx = max(sizes)
        pad = torch.zeros(mx, dtype=torch.float64, device="cuda")
        pad[:len(x)] = torch.from_numpy(np.ascontiguousarray(x)).cuda()
        bufs = [torch.zeros(mx, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(bufs, pad)
        out = [b[:s] for b, s in zip(bufs, sizes)]
    return [o.cpu().numpy() for o in out]


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    nccl_id = exchange_nccl_id(native.nccl_unique_id)
    ctx = native.Context(device=local, rank=rank, nranks=world, nccl_id=nccl_id)
    worst = {}
    case_no = 0
    # float_level 0 (the Poisson options file): ghost values travel as float.  Two correct
    # implementations feed that cast with values that differ by rounding noise (1e-16..1e-13); now
    # and then one of them sits on a float rounding boundary and the cast flips, a 6e-8 relative
    # change of one ghost value.  Over a whole solve that happens with O(1) probability, so with a
    # float halo the residual history can only be compared at ~1e-6; the north_star's 1e-9 is
    # checked on the same partitions with the halo kept in double (use_double forced).
    for name, agg, align, dbl, reb in ((GOLDEN[1], 150, True, False, 0.0), (GOLDEN[1], 0, True, True, 1.02),
                                       (GOLDEN[0], 10 ** 9, True, False, 0.0), (GOLDEN[1], 0, False, True, 0.0),
                                       (GOLDEN[1], 0, False, False, 0.0), (GOLDEN[0], 0, False, True, 0.0)):
        g = Golden(name)
        if dbl:
            for lv in g.hier.levels:
                for op_ in (lv.A, lv.P, lv.R):
                    if op_ is not None:
                        op_.use_double = True
        tol_hist = TOL_HIST if dbl else 1e-6
        hs = partition_hierarchy(g.hier, world, agglomerate_below=agg, align_coarse=align, rebalance_above=reb)
        mine = hs[rank]
        dist.barrier()   # no peer is still writing into the arena this upload is about to replace
        ctx.upload_hierarchy(mine)
        # halo transport: the fused kernel (exchange + SpMV in one launch over NVLink peer memory) on
        # three cases, peer stores with separate launches on one, ncclSend/ncclRecv on one, and one case where the
        # even ranks run the fused kernel and the odd ranks the separate launches on the SAME operators (the two
        # forms speak one hand-shake, csrc/halo_sync.cuh: any mix must work)
        transport = ("fused", "fused", "nccl", "p2p", "fused", "mixed")[case_no % 6]
        if os.environ.get("SAENA_B200_HALO", "p2p") != "p2p":
            transport = "nccl"
        case_no += 1
        use_p2p = transport != "nccl"
        if use_p2p:
            assert setup_p2p_halo(ctx), "peer-memory halo could not be set up"
            ctx.p2p_enable({"fused": 2, "p2p": 1, "mixed": 2 if rank % 2 == 0 else 1}[transport])
            if case_no == 5:
                # per operator: whichever of the two peer-memory paths measures faster (a mix of both)
                ctx.autotune_halo(3)
                transport = "fused/autotuned"
        o = Oracle(hs)
        rng = np.random.default_rng(17)
        tag = (f"{name}/agg{agg}/{'aligned' if align else 'misaligned'}{'+rebalanced' if reb else ''}/{'f64' if dbl else 'f32'}-halo/"
               f"{transport}")
        for l in range(len(mine.levels)):
            sizes = [h.levels[l].A.M for h in hs]
            off = np.concatenate(([0], np.cumsum(sizes)))
            full_v, full_b = rng.standard_normal(off[-1]), rng.standard_normal(off[-1])
            v_parts = [full_v[off[i]:off[i + 1]] for i in range(world)]
            b_parts = [full_b[off[i]:off[i + 1]] for i in range(world)]
            want = o.matvec(l, KIND_A, v_parts)
            got = ctx.matvec(l, KIND_A, v_parts[rank])
            worst[f"{tag}.L{l}.A"] = rel(np.concatenate(gather(got, sizes, rank, world)), np.concatenate(want))
            want = o.smooth(l, "chebyshev", 3, v_parts, b_parts)
            got = ctx.smooth(l, "chebyshev", 3, v_parts[rank], b_parts[rank])
            worst[f"{tag}.L{l}.cheb3"] = rel(np.concatenate(gather(got, sizes, rank, world)), np.concatenate(want))
            if mine.levels[l].P is not None:
                csz = [h.levels[l].P.n_local_cols for h in hs]
                coff = np.concatenate(([0], np.cumsum(csz)))
                full_c = rng.standard_normal(coff[-1])
                c_parts = [full_c[coff[i]:coff[i + 1]] for i in range(world)]
                want = o.matvec(l, KIND_P, c_parts)
                got = ctx.matvec(l, KIND_P, c_parts[rank])
                worst[f"{tag}.L{l}.P"] = rel(np.concatenate(gather(got, sizes, rank, world)), np.concatenate(want))
                want = o.matvec(l, KIND_R, v_parts)
                got = ctx.matvec(l, KIND_R, v_parts[rank])
                worst[f"{tag}.L{l}.R"] = rel(np.concatenate(gather(got, csz, rank, world)), np.concatenate(want))
            if use_p2p and l <= 1:
                # every row mapping of the fused kernel (sliced, sub-warp, row-group), A and P -- and of the
                # separate launches, the streaming mapping included (never fused: it takes the separate launches
                # whatever the mode); then the measurement modes (compute only / exchange only) must leave the
                # hand-shake consistent
                want_a = np.concatenate(o.matvec(l, KIND_A, v_parts))
                for mp in (100, 1, 4, 16, 32, 256, -4, 0):
                    ctx.set_mapping(l, KIND_A, mp)
                    got = ctx.matvec(l, KIND_A, v_parts[rank])
                    worst[f"{tag}.L{l}.A.map{mp}"] = rel(np.concatenate(gather(got, sizes, rank, world)), want_a)
                    if mine.levels[l].P is not None:
                        ctx.set_mapping(l, KIND_P, mp)
                        got = ctx.matvec(l, KIND_P, c_parts[rank])
                        worst[f"{tag}.L{l}.P.map{mp}"] = rel(np.concatenate(gather(got, sizes, rank, world)),
                                                             np.concatenate(o.matvec(l, KIND_P, c_parts)))
                ctx.time_matvec_parts(l, KIND_A, 3)
                got = ctx.matvec(l, KIND_A, v_parts[rank])
                worst[f"{tag}.L{l}.A.after_parts"] = rel(np.concatenate(gather(got, sizes, rank, world)), want_a)
            want = o.vcycle(l, [np.zeros(s) for s in sizes], b_parts)
            got = ctx.vcycle(l, np.zeros(sizes[rank]), b_parts[rank])
            worst[f"{tag}.L{l}.vcycle"] = rel(np.concatenate(gather(got, sizes, rank, world)), np.concatenate(want)) / 10
        sizes = [h.levels[0].A.M for h in hs]
        off = np.concatenate(([0], np.cumsum(sizes)))
        rhs_parts = [g.rhs[off[i]:off[i + 1]] for i in range(world)]
        u_o, it_o, h_o = o.solve_pcg(rhs_parts, g.max_iter, g.tol, "chebyshev", g.pre, g.post)
        ctx.set_graphs(False)
        u_e, it_e, h_e = ctx.solve_pcg(rhs_parts[rank], g.max_iter, g.tol, "chebyshev", g.pre, g.post)
        ctx.set_graphs(True)
        # first V-cycle of the solve runs eagerly, the second is captured (halo flags, peer stores /
        # ncclSend/Recv and the comm stream's fork/join included), the rest replay that graph
        rep0 = ctx.graph_replays()
        u, it, h = ctx.solve_pcg(rhs_parts[rank], g.max_iter, g.tol, "chebyshev", g.pre, g.post)
        replays = ctx.graph_replays() - rep0
        if os.environ.get("SAENA_B200_GRAPH_MULTI", "1") != "0":
            assert replays >= it - 1 >= 1, (tag, "multi-rank V-cycle was not replayed from a graph", replays, it)
        assert it == it_e and np.array_equal(h, h_e) and np.array_equal(u, u_e), (tag, "graph replay != eager solve")
        assert abs(it - it_o) <= 1, (tag, it, it_o)
        n = min(len(h), len(h_o))
        herr = float(np.max(np.abs(h[:n] - h_o[:n]) / h_o[:n]))
        assert herr <= tol_hist, (tag, herr)
        uerr = rel(np.concatenate(gather(u, sizes, rank, world)), np.concatenate(u_o))
        assert uerr < (1e-8 if dbl else 1e-6), (tag, uerr)
        d = ctx.dot(rhs_parts[rank], rhs_parts[rank])
        assert abs(d - float(g.rhs @ g.rhs)) <= 1e-12 * float(g.rhs @ g.rhs)
        if rank == 0:
            print(f"{tag}: pcg iters {it} (oracle {it_o}), history err {herr:.2e}, u err {uerr:.2e}, "
                  f"{replays} V-cycles replayed from a graph", flush=True)
    # ---- failure detection (the counterpart of the reference's print + MPI_Abort, src/saena_object_solve.cpp:1012-1013):
    #      rank 0 applies a distributed operator ALONE.  Its wait for the neighbours' ghost values must run into the
    #      deadline, the call must return an error that names the operator, and after clear_fault + a new import on
    #      every rank the exchange must give the right answer again.
    if os.environ.get("SAENA_B200_HALO", "p2p") == "p2p":
        import time
        sizes = [h.levels[0].A.M for h in hs]
        off = np.concatenate(([0], np.cumsum(sizes)))
        full_v = np.random.default_rng(5).standard_normal(off[-1])
        v_parts = [full_v[off[i]:off[i + 1]] for i in range(world)]
        want = np.concatenate(o.matvec(0, KIND_A, v_parts))
        dist.barrier()
        ctx.set_timeouts(300.0, 60.0)
        if rank == 0:
            t0, msg = time.time(), None
            try:
                ctx.matvec(0, KIND_A, v_parts[0])
            except native.NativeError as e:
                msg = str(e)
            assert msg is not None and "timed out" in msg and "level 0 operator A" in msg, msg
            assert time.time() - t0 < 20.0, "the bounded wait took too long"
            assert ctx.fault_status()
            print(f"bounded wait: a lone application returned after {time.time() - t0:.2f} s with: {msg}", flush=True)
        dist.barrier()
        ctx.clear_fault()
        assert setup_p2p_halo(ctx), "peer-memory halo could not be re-armed after a fault"
        ctx.set_timeouts(5000.0, 180.0)
        assert not ctx.fault_status()
        got = ctx.matvec(0, KIND_A, v_parts[rank])
        worst["after_fault.L0.A"] = rel(np.concatenate(gather(got, sizes, rank, world)), want)
    # ---- the reference's OWN multi-rank layout (tests/golden/*_np{2,4}.npz: per-rank hierarchies exactly as the
    #      reference on `world` MPI ranks laid them out -- shrunk coarse levels, Grid::repart_u plans, float halo --
    #      and its own outputs).  What the drop-in adaptor uploads in a multi-rank run.  Opt-in until its first run
    #      on GPUs (SAENA_MG_REFERENCE_GOLDEN=1): written after the round's GPU budget was spent.
    if os.environ.get("SAENA_MG_REFERENCE_GOLDEN") == "1" and world in (2, 4):
        from tests.util import TOL_HIST_F32_HALO, MultiRankGolden, check_multirank_against_golden
        g = MultiRankGolden({2: "poisson10_np2", 4: "poisson14_np4"}[world])
        mine = g.hiers[rank]
        dist.barrier()
        ctx.upload_hierarchy(mine)
        assert setup_p2p_halo(ctx), "peer-memory halo could not be set up on the reference's layout"
        ctx.autotune_halo(3)

        def apply(name, l, *a):
            sizes = [h.levels[l].A.M for h in g.hiers]
            if name == "A":
                out = ctx.matvec(l, KIND_A, a[0][rank])
            elif name == "P":
                out = ctx.matvec(l, KIND_P, a[0][rank])
            elif name == "R":
                out, sizes = ctx.matvec(l, KIND_R, a[0][rank]), [h.levels[l].R.M for h in g.hiers]
            else:
                out = ctx.smooth(l, "chebyshev" if name == "cheb3" else "jacobi", 3 if name == "cheb3" else 2,
                                 a[0][rank], a[1][rank])
            return gather(out, sizes, rank, world)

        w2 = check_multirank_against_golden(apply, g)
        u, it, h = ctx.solve_pcg(g.rhs[rank], 50, 1e-8, "chebyshev", 3, 3)
        assert it == g.iters, (it, g.iters)
        n = min(len(h), len(g.hist))
        herr = float(np.max(np.abs(h[:n] - g.hist[:n]) / g.hist[:n]))
        assert herr <= TOL_HIST_F32_HALO, herr
        uerr = rel(np.concatenate(gather(u, [len(x) for x in g.u], rank, world)), np.concatenate(g.u))
        assert uerr <= 1e-8, uerr
        if rank == 0:
            print(f"{g.name}: the reference's own {world}-rank layout: worst op err {max(w2.values()):.2e}, pcg iters {it} "
                  f"(reference {g.iters}), history err {herr:.2e}, u err {uerr:.2e}", flush=True)
    bad = {k: e for k, e in worst.items() if not e <= TOL_OP}
    assert not bad, bad
    dist.barrier()
    if rank == 0:
        print(f"MULTIGPU_OK world={world} worst={max(worst.values()):.2e}", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        main()
    except BaseException:
        import traceback
        msg = f"[rank {os.environ.get('RANK')}] FAILED\n" + traceback.format_exc()
        print(msg, flush=True)
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", f"multigpu_fail_rank{os.environ.get('RANK')}.txt"), "w") as f:
                f.write(msg)
        except OSError:
            pass
        os._exit(1)
