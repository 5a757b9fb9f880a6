"""N>1 path on CPU: world_size-2 gloo run of the conditionals broadcast + utterance sharding."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import weights as W
    from chatterbox_b200.dist import broadcast_conditionals, shard_utterances, utterance_cost
    if rank == 0:
        c3, cg = W.make_conds(1234)
    else:                      # other ranks start without the voice
        c3, cg = None, None
    c3, cg = broadcast_conditionals(c3, cg, torch.device("cpu"), src=0)
    ref3, refg = W.make_conds(1234)
    ok = all(torch.equal(c3[k], ref3[k]) for k in ("speaker_emb", "cond_prompt_speech_tokens", "emotion_adv"))
    ok = ok and all(torch.equal(cg[k], refg[k]) for k in ("prompt_token", "prompt_feat", "embedding"))
    ok = ok and cg["prompt_token"].dtype == torch.int64 and int(cg["prompt_token_len"][0]) == 250
    g = torch.Generator().manual_seed(7)
    n_text = torch.randint(16, 160, (64,), generator=g).tolist()
    n_sp = torch.randint(75, 1000, (64,), generator=g).tolist()
    costs = [utterance_cost(a, b) for a, b in zip(n_text, n_sp)]
    shards = shard_utterances(costs, world)
    mine = torch.tensor(sorted(shards[rank]))
    gathered = [torch.zeros(64, dtype=torch.long) for _ in range(world)]
    pad = torch.full((64,), -1, dtype=torch.long)
    pad[:len(mine)] = mine
    dist.all_gather(gathered, pad)
    allidx = sorted(int(i) for t in gathered for i in t.tolist() if i >= 0)
    ok = ok and allidx == list(range(64))
    loads = [sum(costs[i] for i in s) for s in shards]
    ok = ok and max(loads) / (sum(loads) / world) < 1.1           # LPT balance
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_and_sharding_world2_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + os.getpid() % 500
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1], dict(out)
