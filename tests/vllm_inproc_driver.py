"""Boot a real vLLM engine (tiny Llama, REAL random weights, no tokenizer) with a KV connector in
the plugin slot and run the same prompts twice: the first pass computes every prompt and stores its
KV through the connector, the second pass finds the KV in the pinned pool and loads it instead of
prefilling.  Prints one JSON line with both passes' greedy tokens and top-k logprobs and the pool /
connector counters.  Started as a subprocess by tests/test_gpu_vllm_connector.py (never imported by
the product).

The slot is the one production-stack fills (helm/templates/deployment-vllm-multi.yaml:194-207);
`--connector native` loads B200KVConnector by module path, `--connector alias` names the chart's
literal `LMCacheConnectorV1`, which vLLM resolves to its own wrapper
(vllm/.../kv_connector/v1/lmcache_connector.py:105-113) and which imports
`lmcache.integration.vllm.vllm_v1_adapter.LMCacheConnectorV1Impl` — here this repo's compat tree.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_model(path: str, layers: int, vocab: int, seed: int = 0) -> None:
    """A Llama with the KV geometry of Llama-3-8B per layer (8 KV heads x 128) and real N(0, 0.02)
    weights, so logits depend on the KV that is read (vLLM's --load-format dummy gives near-constant
    logits, useless for an FP8 tolerance)."""
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(seed)
    cfg = LlamaConfig(hidden_size=1024, intermediate_size=2816, num_hidden_layers=layers, num_attention_heads=8,
                      num_key_value_heads=8, head_dim=128, vocab_size=vocab, max_position_embeddings=4096,
                      rms_norm_eps=1e-5, rope_theta=500000.0, tie_word_embeddings=False, bos_token_id=1,
                      eos_token_id=2, initializer_range=0.04, torch_dtype="bfloat16")
    m = LlamaForCausalLM(cfg).to(torch.bfloat16)
    m.save_pretrained(path, safe_serialization=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--connector", default="native", choices=["native", "alias", "none"])
    ap.add_argument("--compiled", action="store_true", help="default vLLM mode (torch.compile + CUDA graphs) instead of eager")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--vocab", type=int, default=4096)
    ap.add_argument("--model-dir", default="/tmp/b200kv-tiny-llama")
    ap.add_argument("--max-tokens", type=int, default=12)
    ap.add_argument("--out", default="")
    a = ap.parse_args()

    if not os.path.exists(os.path.join(a.model_dir, "config.json")):
        make_model(a.model_dir, a.layers, a.vocab)

    import numpy as np
    from vllm import LLM, SamplingParams
    from vllm.config import KVTransferConfig

    ktc = None
    if a.connector == "native":
        ktc = KVTransferConfig(kv_connector="B200KVConnector", kv_connector_module_path="b200kv.connector",
                               kv_role="kv_both", kv_load_failure_policy="recompute")
    elif a.connector == "alias":
        ktc = KVTransferConfig(kv_connector="LMCacheConnectorV1", kv_role="kv_both", kv_load_failure_policy="recompute")
    kw = {}
    if a.connector == "alias":
        kw["disable_hybrid_kv_cache_manager"] = True     # the wrapper is not SupportsHMA
    llm = LLM(model=a.model_dir, skip_tokenizer_init=True, dtype="bfloat16", seed=0, max_model_len=4096,
              gpu_memory_utilization=0.25, enable_prefix_caching=False, enforce_eager=not a.compiled,
              kv_transfer_config=ktc, **kw)

    rng = np.random.default_rng(1234)
    shared = rng.integers(3, a.vocab, 512).tolist()
    prompts = []
    for i, n in enumerate((300, 700, 1029, 1536, 2100, 2817)):
        body = rng.integers(3, a.vocab, n).tolist()
        prompts.append((shared + body) if i % 2 else body)       # odd ones share a 512-token (2-chunk) prefix
    sp = SamplingParams(temperature=0.0, max_tokens=a.max_tokens, logprobs=5, detokenize=False, ignore_eos=True)

    def run():
        outs = llm.generate([{"prompt_token_ids": p} for p in prompts], sp, use_tqdm=False)
        res = []
        for o in outs:
            c = o.outputs[0]
            steps = []
            for t, step in zip(c.token_ids, c.logprobs or []):
                top = sorted((float(v.logprob) for v in step.values()), reverse=True)
                steps.append({"lp": float(step[t].logprob), "margin": top[0] - top[1] if len(top) > 1 else 99.0})
            res.append({"tokens": [int(t) for t in c.token_ids], "steps": steps,
                        "num_cached_tokens": getattr(o, "num_cached_tokens", None)})
        return res

    first = run()
    time.sleep(1.0)          # let the last D2H commits land (a store never blocks the forward pass)
    second = run()

    metrics = {}
    try:
        for m in llm.get_metrics():
            if any(s in m.name for s in ("external_prefix_cache", "lmcache", "b200kv")):
                v = getattr(m, "value", None)
                if v is not None:
                    metrics[m.name] = metrics.get(m.name, 0) + float(v)
    except Exception as e:  # metrics are a bonus, the pool counters below are the evidence
        metrics["error"] = repr(e)
    out = {"connector": a.connector, "compiled": a.compiled, "format": os.environ.get("B200KV_FORMAT", "raw"),
           "prompt_lens": [len(p) for p in prompts], "first": first, "second": second, "metrics": metrics}
    line = json.dumps(out)
    if a.out:
        with open(a.out, "w") as f:
            f.write(line)
    print("RESULT " + line, flush=True)


if __name__ == "__main__":
    sys.exit(main())
