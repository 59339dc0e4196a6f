"""utils/Config.scala:3-21 + resources/application.conf:1-52 -- the `dsgd { ... }` parameter contract.

Same keys, same defaults, same DSGD_* environment overrides.  The reference resolves them with
pureconfig over HOCON; here a minimal reader handles the subset of HOCON that application.conf uses
(`key = value` lines inside `dsgd { }`, `${?ENV}` optional substitutions, `#` comments).
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, fields
from typing import Dict, Optional


@dataclass
class Config:  # field order and names: utils/Config.scala:3-21 (kebab-case in the file)
    host: str = "127.0.0.1"
    port: int = 4000
    master_host: Optional[str] = None
    master_port: Optional[int] = None
    batch_size: int = 100
    learning_rate: float = 0.5
    lam: float = 0.00001          # `lambda`
    node_count: int = 3
    full: bool = False
    is_async: bool = False        # `async`
    record: bool = False
    data_path: str = "data"
    max_epochs: int = 10
    check_every: int = 100
    leaky_loss: float = 0.9
    conv_delta: float = 0.01
    patience: int = 5


# application.conf key -> (Config field, DSGD_* variable)   (resources/application.conf:2-50)
_KEYS = {
    "data-path": ("data_path", "DSGD_DATA_PATH"), "host": ("host", "DSGD_NODE_HOST"), "port": ("port", "DSGD_NODE_PORT"),
    "master-host": ("master_host", "DSGD_MASTER_HOST"), "master-port": ("master_port", "DSGD_MASTER_PORT"),
    "batch-size": ("batch_size", "DSGD_BATCH_SIZE"), "learning-rate": ("learning_rate", "DSGD_LEARNING_RATE"),
    "lambda": ("lam", "DSGD_LAMBDA"), "full": ("full", "DSGD_FULL"), "node-count": ("node_count", "DSGD_NODE_COUNT"),
    "async": ("is_async", "DSGD_ASYNC"), "record": ("record", "DSGD_RECORD"), "max-epochs": ("max_epochs", "DSGD_MAX_EPOCHS"),
    "check-every": ("check_every", "DSGD_CHECK_EVERY"), "leaky-loss": ("leaky_loss", "DSGD_LEAKY_LOSS"),
    "patience": ("patience", "DSGD_PATIENCE"), "conv-delta": ("conv_delta", "DSGD_CONV_DELTA"),
}
_TYPES = {f.name: f.type for f in fields(Config)}


def _coerce(field: str, raw: str):
    raw = raw.strip().strip('"')
    t = str(_TYPES[field])
    if "bool" in t:
        if raw.lower() in ("true", "yes", "on"):
            return True
        if raw.lower() in ("false", "no", "off"):
            return False
        raise ValueError(f"{field}: expected a boolean, got {raw!r}")
    if "int" in t:
        return int(raw)
    if "float" in t:
        return float(raw)
    return raw


def _parse_block(text: str, env: Dict[str, str]) -> Dict[str, str]:
    """Returns key -> raw value for the `dsgd { }` block; later assignments win, `${?X}` only if X is set."""
    m = re.search(r"\bdsgd\s*\{", text)
    if not m:
        return {}
    depth, i = 1, m.end()
    while i < len(text) and depth:
        depth += {"{": 1, "}": -1}.get(text[i], 0)
        i += 1
    out: Dict[str, str] = {}
    for line in text[m.end():i - 1].splitlines():
        line = line.split("#", 1)[0].strip()
        if not line or "=" not in line:
            continue
        key, raw = (s.strip() for s in line.split("=", 1))
        sub = re.fullmatch(r"\$\{\?(\w+)\}", raw)
        if sub:
            if sub.group(1) in env:
                out[key] = env[sub.group(1)]
            continue
        out[key] = raw
    return out


def load_config(path: Optional[str] = None, env: Optional[Dict[str, str]] = None) -> Config:
    """`pureconfig.loadConfigOrThrow[Config]("dsgd")` (Main.scala:36).  Without a file the defaults of
    resources/application.conf apply; DSGD_* variables override either."""
    env = dict(os.environ if env is None else env)
    cfg = Config()
    raw: Dict[str, str] = {}
    if path is not None:
        with open(path) as f:
            raw = _parse_block(f.read(), env)
    else:
        for key, (_, var) in _KEYS.items():
            if var in env:
                raw[key] = env[var]
    for key, value in raw.items():
        if key not in _KEYS:
            raise KeyError(f"unknown key dsgd.{key}")
        field = _KEYS[key][0]
        setattr(cfg, field, _coerce(field, value))
    return cfg
