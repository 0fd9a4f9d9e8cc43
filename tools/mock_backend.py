"""Minimal OpenAI-compatible mock engine for router plumbing tests (BASELINE.json configs[0]):
/v1/models, /v1/completions, /v1/chat/completions (optionally streamed), /tokenize, /health,
/metrics.  Own implementation — the reference's src/tests/perftest/fake-openai-server.py imports
vLLM protocol modules that no longer exist in vLLM 0.22.

    python tools/mock_backend.py --host 127.0.0.1 --port 9001 --model m
"""
import argparse
import asyncio
import hashlib
import json
import time

import uvicorn
from fastapi import FastAPI, Request
from fastapi.responses import JSONResponse, PlainTextResponse, StreamingResponse

app = FastAPI()
STATE = {"model": "m", "served": [], "name": "backend"}


@app.get("/health")
async def health():
    return PlainTextResponse("ok")


@app.get("/v1/models")
async def models():
    return {"object": "list", "data": [{"id": STATE["model"], "object": "model", "created": 0, "owned_by": "mock"}]}


@app.get("/metrics")
async def metrics():
    n = len(STATE["served"])
    return PlainTextResponse(
        f'vllm:num_requests_running{{model_name="{STATE["model"]}"}} 0\n'
        f'vllm:num_requests_waiting{{model_name="{STATE["model"]}"}} 0\n'
        f'vllm:gpu_cache_usage_perc{{model_name="{STATE["model"]}"}} 0.0\n'
        f'vllm:gpu_prefix_cache_hit_rate{{model_name="{STATE["model"]}"}} 0.0\n'
        f'mock_requests_total {n}\n')


@app.get("/served")
async def served():
    return {"name": STATE["name"], "served": STATE["served"]}


@app.post("/tokenize")
async def tokenize(req: Request):
    body = await req.json()
    toks = [hash(w) % 50000 for w in str(body.get("prompt", "")).split()]
    return {"tokens": toks, "count": len(toks), "max_model_len": 4096}


def _record(req: Request, body: dict):
    full = str(body.get("prompt", ""))
    STATE["served"].append({"id": req.headers.get("x-request-id"), "user": req.headers.get("x-user-id"),
                            "t": time.time(), "prompt": full[:64], "sha1": hashlib.sha1(full.encode()).hexdigest(),
                            "n_messages": body.get("n_messages"), "max_tokens": body.get("max_tokens"),
                            "stream": bool(body.get("stream"))})


@app.post("/v1/completions")
async def completions(req: Request):
    body = await req.json()
    _record(req, body)
    n = int(body.get("max_tokens", 8))
    text = " ".join(["tok"] * n)
    usage = {"prompt_tokens": len(str(body.get("prompt", "")).split()), "completion_tokens": n,
             "total_tokens": n + len(str(body.get("prompt", "")).split())}
    if body.get("stream"):
        async def gen():
            for i in range(n):
                chunk = {"id": "cmpl-mock", "object": "text_completion", "model": STATE["model"],
                         "choices": [{"index": 0, "text": "tok ", "finish_reason": None}]}
                yield f"data: {json.dumps(chunk)}\n\n"
                await asyncio.sleep(0)
            yield "data: [DONE]\n\n"
        return StreamingResponse(gen(), media_type="text/event-stream")
    return JSONResponse({"id": "cmpl-mock", "object": "text_completion", "created": int(time.time()),
                         "model": STATE["model"], "served_by": STATE["name"],
                         "choices": [{"index": 0, "text": text, "finish_reason": "length"}], "usage": usage})


@app.post("/v1/chat/completions")
async def chat(req: Request):
    body = await req.json()
    _record(req, {"prompt": json.dumps(body.get("messages", [])), "n_messages": len(body.get("messages", [])),
                  "max_tokens": body.get("max_tokens"), "stream": body.get("stream")})
    n = int(body.get("max_tokens", 8))
    if body.get("stream"):
        async def gen():
            for i in range(n):
                chunk = {"id": "chat-mock", "object": "chat.completion.chunk", "model": STATE["model"],
                         "choices": [{"index": 0, "delta": {"content": "tok "}, "finish_reason": None}]}
                yield f"data: {json.dumps(chunk)}\n\n"
                await asyncio.sleep(0)
            yield "data: " + json.dumps({"id": "chat-mock", "object": "chat.completion.chunk", "choices": [],
                                         "usage": {"prompt_tokens": 1, "completion_tokens": n, "total_tokens": n + 1}}) + "\n\n"
            yield "data: [DONE]\n\n"
        return StreamingResponse(gen(), media_type="text/event-stream")
    return {"id": "chat-mock", "object": "chat.completion", "model": STATE["model"],
            "choices": [{"index": 0, "message": {"role": "assistant", "content": " ".join(["tok"] * n)},
                         "finish_reason": "length"}],
            "usage": {"prompt_tokens": 1, "completion_tokens": n, "total_tokens": n + 1}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, required=True)
    ap.add_argument("--model", default="m")
    ap.add_argument("--name", default=None)
    a = ap.parse_args()
    STATE["model"] = a.model
    STATE["name"] = a.name or f"{a.host}:{a.port}"
    uvicorn.run(app, host=a.host, port=a.port, log_level="warning")
