"""Orchestration messages the router constructs and the replies it reads
(src/vllm_router/routers/routing_logic.py:378-387 LookupMsg / layout_info,
:413-423 QueryInstMsg / instance_id)."""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass
class LookupMsg:
    tokens: list
    event_id: str = ""


@dataclass
class LookupRetMsg:
    # instance_id -> (location, matched_tokens); the router reads [1] of the FIRST key
    layout_info: dict = field(default_factory=dict)
    event_id: str = ""


@dataclass
class QueryInstMsg:
    ip: str
    event_id: str = ""


@dataclass
class QueryInstRetMsg:
    instance_id: str | None
    event_id: str = ""


# ---- worker -> controller (this shim's own wire format: JSON over ZMQ PUSH/PULL) -------------
@dataclass
class RegisterMsg:
    instance_id: str
    ip: str
    pool_name: str       # POSIX shm segment holding the chunk index (same host)
    key_seed: int
    chunk_tokens: int
    owner_tag: int = 0
    include_partial: bool = True


@dataclass
class HeartbeatMsg:
    instance_id: str
