"""Device-side callers of the step path: scripted opponents, masked categorical
sampling of policy logits and GAE -- thin wrappers over spl_scripted_action / spl_masked_sample / spl_gae."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib as L

BOTS = {"random": L.BOT_RANDOM, "greedy_v1": L.BOT_GREEDY_V1, "basic": L.BOT_BASIC_PRIORITY, "basic_priority": L.BOT_BASIC_PRIORITY,
        "greedy_v2": L.BOT_GREEDY_V2}


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def scripted_action(obs: torch.Tensor, mask: torch.Tensor, kind: str = "greedy_v1", *, key: int = 0xB07, t: int = 0, env_offset: int = 0,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One action per env from a scripted opponent of the reference (scripts/eval_suite.py:10-128; wrappers/selfplay.py:66-73).
    Usable directly as ``opponent_policy`` of ``SplendorVecEnv.dual_step`` via ``bot_policy(kind)``."""
    n = mask.shape[0]
    assert obs.dtype == torch.int32 and mask.dtype == torch.int8 and obs.is_contiguous() and mask.is_contiguous()
    out = torch.empty(n, dtype=torch.int32, device=mask.device) if out is None else out
    with torch.cuda.device(mask.device):
        L.check(L.load().spl_scripted_action(obs.data_ptr(), mask.data_ptr(), n, BOTS[kind], env_offset, key, t, out.data_ptr(), _stream(mask)),
                "spl_scripted_action")
    return out


def bot_policy(kind: str, key: int = 0xB07):
    """``opponent_policy(obs, mask) -> actions`` closure with its own step counter for the random tie-breaks."""
    state = {"t": 0}

    def policy(obs, mask):
        state["t"] += 1
        return scripted_action(obs, mask, kind, key=key, t=state["t"])

    return policy


def masked_sample(logits: torch.Tensor, mask: torch.Tensor, *, greedy: bool = False, key: int = 0x5A11, t: int = 0, env_offset: int = 0,
                  want_logprob: bool = True, want_entropy: bool = False) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Masked categorical over ``logits [N,45]`` (float32, or float16 with any row pitch): sample (ppo_splendor.py:27-38,54-59) or argmax
    (scripts/eval_suite.py:131-141).  Returns (actions int32, log_prob, entropy)."""
    n = mask.shape[0]
    actions = torch.empty(n, dtype=torch.int32, device=mask.device)
    logprob = torch.empty(n, dtype=torch.float32, device=mask.device) if want_logprob else None
    entropy = torch.empty(n, dtype=torch.float32, device=mask.device) if want_entropy else None
    lp, en = (None if logprob is None else logprob.data_ptr()), (None if entropy is None else entropy.data_ptr())
    with torch.cuda.device(mask.device):
        if logits.dtype == torch.float16 and logits.dim() == 2 and logits.stride(1) == 1 and logits.shape[1] >= L.NUM_ACTIONS:
            # half-precision head, possibly a column slice of a padded output ([:, :45] of [N, 48]): read in place
            L.check(L.load().spl_masked_sample_f16(logits.data_ptr(), logits.stride(0), mask.data_ptr(), n, 1 if greedy else 0, env_offset,
                                                   key, t, actions.data_ptr(), lp, en, _stream(mask)), "spl_masked_sample_f16")
        else:
            logits = logits.float().contiguous()
            L.check(L.load().spl_masked_sample(logits.data_ptr(), mask.data_ptr(), n, 1 if greedy else 0, env_offset, key, t,
                                               actions.data_ptr(), lp, en, _stream(mask)), "spl_masked_sample")
    return actions, logprob, entropy


def gae(rewards: torch.Tensor, values: torch.Tensor, terminals: torch.Tensor, last_values: torch.Tensor, gamma: float = 0.99,
        gae_lambda: float = 0.95) -> Tuple[torch.Tensor, torch.Tensor]:
    """Advantages and returns over step-major ``[T,N]`` buffers (ppo_splendor.py:299-314)."""
    T, n = rewards.shape
    rewards, values = rewards.float().contiguous(), values.float().contiguous()
    terminals = terminals.view(torch.uint8) if terminals.dtype == torch.bool else terminals.to(torch.uint8)
    terminals, last_values = terminals.contiguous(), last_values.float().contiguous()
    adv, ret = torch.empty_like(rewards), torch.empty_like(rewards)
    with torch.cuda.device(rewards.device):
        L.check(L.load().spl_gae(rewards.data_ptr(), values.data_ptr(), terminals.data_ptr(), last_values.data_ptr(), T, n, gamma, gae_lambda,
                                 adv.data_ptr(), ret.data_ptr(), _stream(rewards)), "spl_gae")
    return adv, ret
