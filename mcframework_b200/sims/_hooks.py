"""Shared plumbing of the built-in hooks: one replication on the device from a live NumPy Generator."""
from __future__ import annotations

import numpy as np

from .. import _device as D
from .. import _streams as S


def one_replication(prog, rng: np.random.Generator) -> float:
    """``single_simulation`` called directly: replay ONE replication from ``rng``'s current position on the
    device and leave ``rng`` where NumPy would have left it."""
    kind, rec = S.generator_record(rng, 0, 1)
    dl = D.DeviceLayout.of(S.StreamLayout(kind, rec, 1, 0))
    values, used = prog.execute(dl)
    draws = used if isinstance(used, int) else int(used.cpu().numpy()[0])
    value = float(values.reshape(-1)[0].item())
    S.consume_draws(rng, draws)
    return value
