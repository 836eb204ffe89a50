"""Small host-side helpers shared by the launcher paths (no torch, no CUDA)."""
from __future__ import annotations

import socket


def free_port() -> int:
    """A TCP port that was free a moment ago on loopback (rendezvous ports of locally started jobs)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p
