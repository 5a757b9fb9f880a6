"""Multi-GPU plumbing: one process per GPU, utterances sharded across ranks, NCCL only for the broadcast of the
voice conditionals (the north_star's "speaker-embedding broadcast").  No collective on the hot path:
utterances are independent (the reference itself is batch-1), so every rank decodes / solves / vocodes its own shard.
The same code runs on CPU with the gloo backend in tests/test_dist_cpu.py."""
import numpy as np
import torch
import torch.distributed as dist

_FIELDS = [("t3", "speaker_emb", torch.float32), ("t3", "cond_prompt_speech_tokens", torch.int64),
           ("t3", "emotion_adv", torch.float32), ("gen", "prompt_token", torch.int64),
           ("gen", "prompt_feat", torch.float32), ("gen", "embedding", torch.float32)]


def pack_conditionals(t3_cond: dict, gen: dict):
    """Flatten the per-voice conditionals (~165 KB) into one fp32 buffer + the shapes needed to unpack."""
    parts, meta = [], []
    for grp, key, dt in _FIELDS:
        t = (t3_cond if grp == "t3" else gen)[key]
        t = torch.as_tensor(t)
        meta.append((grp, key, tuple(t.shape), dt))
        parts.append(t.reshape(-1).to(torch.float64 if dt == torch.int64 else torch.float32).to(torch.float32))
    return torch.cat(parts), meta


def unpack_conditionals(flat, meta):
    t3, gen, o = {}, {}, 0
    for grp, key, shape, dt in meta:
        n = int(np.prod(shape)) if len(shape) else 1
        t = flat[o:o + n].reshape(shape)
        t = t.round().to(torch.int64) if dt == torch.int64 else t.to(torch.float32)
        (t3 if grp == "t3" else gen)[key] = t
        o += n
    gen["prompt_token_len"] = torch.tensor([gen["prompt_token"].shape[-1]])
    gen["prompt_feat_len"] = None
    return t3, gen


def broadcast_conditionals(t3_cond, gen, device, src=0):
    """Rank `src` owns the voice; everybody else receives it.  Token ids (< 6561) are exact in fp32."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t3_cond, gen
    rank = dist.get_rank()
    if rank == src:
        flat, meta = pack_conditionals(t3_cond, gen)
        obj = [meta]
    else:
        flat, obj = None, [None]
    dist.broadcast_object_list(obj, src=src)           # tiny metadata (shapes)
    meta = obj[0]
    n = sum(int(np.prod(m[2])) if len(m[2]) else 1 for m in meta)
    buf = flat.to(device) if rank == src else torch.zeros(n, dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src)                       # the payload: NCCL over NVLink on the GPU box
    return unpack_conditionals(buf.cpu(), meta)


def shard_utterances(costs, world):
    """Longest-processing-time-first assignment of utterances to ranks (strong-scaling mode).  Returns a list of
    index lists, one per rank; every utterance appears exactly once."""
    order = np.argsort(-np.asarray(costs, dtype=np.float64), kind="stable")
    loads = np.zeros(world)
    shards = [[] for _ in range(world)]
    for i in order:
        r = int(np.argmin(loads))
        shards[r].append(int(i))
        loads[r] += float(costs[i])
    return shards


def utterance_cost(n_text, n_speech, n_prompt=250):
    """Relative cost model: KV traffic of the AR decode + attention-dominated CFM."""
    s0 = 34 + n_text + 2
    t = 2.0 * (n_prompt + n_speech)
    return n_speech * (s0 + 0.5 * n_speech) + 0.02 * t * t
