"""Greedy-output equivalence (SURVEY.md §8c end-to-end criterion): temperature-0 answers of the
same multi-round conversations served with the connector (KV loaded from the pool on later turns)
vs without it (recomputed).  Compares the per-request SHA-1 of the generated text.
    python tools/e2e/compare_outputs.py gpurun_out/mrqa_rows_none.jsonl gpurun_out/mrqa_rows_b200kv.jsonl"""
import json
import sys

a = {(r["user_id"], r["question_id"]): r for r in map(json.loads, open(sys.argv[1])) if "error" not in r}
b = {(r["user_id"], r["question_id"]): r for r in map(json.loads, open(sys.argv[2])) if "error" not in r}
keys = sorted(set(a) & set(b))
by_turn = {}
for k in keys:
    t = by_turn.setdefault(k[1], [0, 0])
    t[1] += 1
    t[0] += a[k]["body_sha1"] == b[k]["body_sha1"]
# a conversation stays comparable only while every earlier answer matched (answers feed later prompts)
chain_ok, chain_n = 0, 0
for u in sorted({k[0] for k in keys}):
    alive = True
    for q in sorted(k[1] for k in keys if k[0] == u):
        if not alive:
            break
        chain_n += 1
        alive = a[(u, q)]["body_sha1"] == b[(u, q)]["body_sha1"]
        chain_ok += alive
print(json.dumps({"requests_compared": len(keys), "identical_by_turn": {t: f"{v[0]}/{v[1]}" for t, v in sorted(by_turn.items())},
                  "identical_while_history_identical": f"{chain_ok}/{chain_n}"}))
