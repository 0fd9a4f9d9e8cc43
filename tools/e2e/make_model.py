"""Synthesize an offline Llama-3-8B-shaped model directory (no weights, no network):
config.json with the real architecture numbers (32 layers, 32 heads, 8 KV heads, head 128, hidden
4096, vocab 128256 — SURVEY.md §7), a word-level tokenizer in which the harness's dummy word "hi"
is one token (benchmarks/multi-round-qa/multi-round-qa.py:234-243 builds prompts from "hi" x N),
and a minimal chat template.  Used with `vllm serve <dir> --load-format dummy`.

    python tools/e2e/make_model.py /tmp/llama3-8b-synth [--layers 32]
"""
import argparse
import json
import os

from tokenizers import Tokenizer, models, pre_tokenizers, processors


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--layers", type=int, default=32)
    ap.add_argument("--max-len", type=int, default=8192)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    cfg = {
        "architectures": ["LlamaForCausalLM"], "model_type": "llama", "hidden_size": 4096,
        "intermediate_size": 14336, "num_hidden_layers": a.layers, "num_attention_heads": 32,
        "num_key_value_heads": 8, "head_dim": 128, "vocab_size": 128256, "max_position_embeddings": a.max_len,
        "rms_norm_eps": 1e-5, "rope_theta": 500000.0, "hidden_act": "silu", "tie_word_embeddings": False,
        "torch_dtype": "bfloat16", "bos_token_id": 1, "eos_token_id": 2, "attention_bias": False,
        "mlp_bias": False, "use_cache": True,
    }
    json.dump(cfg, open(os.path.join(a.out, "config.json"), "w"), indent=1)
    json.dump({"bos_token_id": 1, "eos_token_id": 2, "do_sample": False},
              open(os.path.join(a.out, "generation_config.json"), "w"))
    words = ["<unk>", "<s>", "</s>", "<|user|>", "<|assistant|>", "<|system|>", "hi", "Hi", "here", "s", "some",
             "system", "prompt", "For", "user", "are", "other", "context", "Here", "question", "can", "you", "tell",
             "me", "a", "new", "long", "story", "with", "happy", "ending", ".", ",", ":", "?", "#", "'", "-"]
    words += [str(i) for i in range(10)]
    words += [chr(c) for c in range(ord("a"), ord("z") + 1) if chr(c) not in words]
    words += [chr(c) for c in range(ord("A"), ord("Z") + 1) if chr(c) not in words]
    words = list(dict.fromkeys(words))
    # fill the whole model vocabulary so every id a random-init model samples decodes to text
    words += [f"w{i}" for i in range(len(words), cfg["vocab_size"])]
    vocab = {w: i for i, w in enumerate(words)}
    tok = Tokenizer(models.WordLevel(vocab=vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Sequence([pre_tokenizers.WhitespaceSplit(), pre_tokenizers.Punctuation(),
                                                 pre_tokenizers.Digits(individual_digits=True)])
    tok.post_processor = processors.TemplateProcessing(single="<s> $A", special_tokens=[("<s>", 1)])
    tok.save(os.path.join(a.out, "tokenizer.json"))
    tmpl = ("{% for m in messages %}{% if m['role'] == 'user' %}<|user|> {{ m['content'] }} "
            "{% elif m['role'] == 'system' %}<|system|> {{ m['content'] }} "
            "{% else %}<|assistant|> {{ m['content'] }} {% endif %}{% endfor %}"
            "{% if add_generation_prompt %}<|assistant|>{% endif %}")
    json.dump({"tokenizer_class": "PreTrainedTokenizerFast", "bos_token": "<s>", "eos_token": "</s>",
               "unk_token": "<unk>", "model_max_length": a.max_len, "chat_template": tmpl,
               "clean_up_tokenization_spaces": False},
              open(os.path.join(a.out, "tokenizer_config.json"), "w"), indent=1)
    json.dump({"bos_token": "<s>", "eos_token": "</s>", "unk_token": "<unk>"},
              open(os.path.join(a.out, "special_tokens_map.json"), "w"))
    print("wrote", a.out)


if __name__ == "__main__":
    main()
