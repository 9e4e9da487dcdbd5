"""Worker of tests/test_ddp_nccl_gpu.py (launched with torch.distributed.run, one rank per GPU; not collected by pytest).

N NCCL ranks, each stepping on its slice of a global batch through `UnetFineTuner.train_step` (per-block bucketed all-reduce
overlapped with backward, AdamW inside the backward) must equal ONE rank stepping on the whole batch -- the contract of the
reference's DDP wrap (pdm/training/trainer.py:122-129,2257-2260; SURVEY section 4 item iv)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    solo = [dist.new_group([r]) for r in range(world)][rank]          # a group of one: the single-rank reference run
    from oracle import pdm_restated as P
    from oracle.make_golden import SMALL64, deterministic_fill, make_arch_vector
    from unlearn_ft_b200.pdm.models import UNet2DConditionModel, UNet2DConditionModelPruned
    from unlearn_ft_b200.pdm.training import UnetFineTuner
    cfg = dict(block_out_channels=SMALL64["block_out_channels"], attention_head_dim=SMALL64["heads"],
               cross_attention_dim=SMALL64["cross_attention_dim"])
    full = P.UNetGated(**SMALL64)
    deterministic_fill(full, 3)
    av = make_arch_vector(full.get_structure(), 0.55, 21, (2,))
    teacher = UNet2DConditionModel(cfg, seed=7)

    def student():
        m = UNet2DConditionModelPruned(cfg, arch_vector=av, seed=None)
        m.load_unpruned_state_dict(full.state_dict())
        return m

    per = 2
    B = per * world
    g = torch.Generator().manual_seed(0)
    glob = dict(latents=torch.randn(B, 4, 16, 16, generator=g).cuda(), noise=torch.randn(B, 4, 16, 16, generator=g).cuda(),
                timesteps=torch.randint(0, 1000, (B,), generator=g).cuda(),
                prompt_embeds=torch.randn(B, 77, SMALL64["cross_attention_dim"], generator=g).cuda())
    mine = {k: v[rank * per:(rank + 1) * per].contiguous() for k, v in glob.items()}
    results = {}
    for mode in ("eager", "graph"):
        ddp = UnetFineTuner(student(), teacher, lr=1e-4, warmup_steps=0)
        assert ddp.reducer.world == world
        local = UnetFineTuner(student(), teacher, lr=1e-4, warmup_steps=0, process_group=solo)   # same slice, no exchange
        ref = UnetFineTuner(student(), teacher, lr=1e-4, warmup_steps=0, process_group=solo)     # the whole batch on one rank
        assert ref.reducer.world == 1 and local.reducer.world == 1
        if mode == "graph":
            ddp.capture_cuda_graph(mine)
        l_ddp = torch.stack([v.detach().float() for v in ddp.train_step(mine)]).clone()
        local.train_step(mine)
        l_ref = torch.stack([v.detach().float() for v in ref.train_step(glob)])
        dist.all_reduce(l_ddp, op=dist.ReduceOp.AVG)                  # mean over the global batch = mean of the rank means
        assert ddp.reducer.shards, "the optimizer update is expected to be sharded over the ranks"
        own = sum(s for s in ddp.reducer.shards.values())
        assert abs(own * world - sum(hi - lo for lo, hi in ddp.reducer.shards)) == 0
        ddp.reducer.gather_optimizer_state(ddp.optimizer)             # moments live on the slice owners until gathered
        plain = UnetFineTuner(student(), teacher, lr=1e-4, warmup_steps=0)
        plain.reducer.shard = False                                   # replicated update through plain all-reduces
        plain.train_step(mine)
        torch.cuda.synchronize()
        shard_err = ((ddp.student.arena.master.detach().double() - plain.student.arena.master.detach().double()).abs().max()).item()
        assert not plain.reducer.shards and shard_err <= 2.1e-4, shard_err   # (Adam's first step moves every weight by +-lr = 1e-4)
        assert torch.equal(ddp.student.arena.shadow, ddp.student.arena.master.detach().bfloat16())
        assert float(ddp.student.arena.grad.abs().sum()) == 0.0
        # first-step AdamW moment = (1 - beta1) * gradient: compares the exchanged gradient itself
        m_ddp, m_ref = ddp.optimizer.exp_avg.double(), ref.optimizer.exp_avg.double()
        m_avg = local.optimizer.exp_avg.clone()
        dist.all_reduce(m_avg, op=dist.ReduceOp.AVG)                  # torch's own average of the ranks' local gradients
        comm_err = ((m_ddp - m_avg.double()).norm() / m_avg.double().norm()).item()
        grad_err = ((m_ddp - m_ref).norm() / m_ref.norm()).item()
        loss_err = ((l_ddp - l_ref).abs() / l_ref.abs()).max().item()
        # every rank holds the same parameters afterwards
        chk = ddp.student.arena.master.detach().double().sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN), dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        results[mode] = (comm_err, grad_err, loss_err, float(hi - lo))
        print(f"[rank {rank}] {mode}: exchanged gradient vs mean of local gradients rel-L2 {comm_err:.2e}; vs one rank on the "
              f"concatenated batch: gradient rel-L2 {grad_err:.2e}, loss rel {loss_err:.2e}; replica spread {float(hi - lo):.1e}",
              flush=True)
        ddp.release_cuda_graph()
    for mode, (ce, ge, le, spread) in results.items():
        # The forward pass is bit-reproducible and per-sample arithmetic does not depend on the batch size, so N ranks on slices
        # of a batch and ONE rank on the whole batch differ only by fp32 summation order in the weight-gradient reductions and
        # the collective: measured 5e-8 (gradient, relative L2) and 0 (losses).
        assert ce < 1e-5 and ge < 1e-5 and le < 1e-6 and spread == 0.0, (mode, ce, ge, le, spread)
    print(f"[rank {rank}] ddp nccl parity ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
