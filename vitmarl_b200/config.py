"""Static configuration mirroring the reference's frozen dataclasses
(``gymnax_exchange/jaxob/jaxob_config.py:11-29,158-183`` and ``jaxob_constants.py``).
Only the fields the hot path reads are kept; names and defaults are the reference's."""
from __future__ import annotations

from dataclasses import dataclass

MAXINT = 2_147_483_647     # jaxob_constants.py:4-6
INITID = -2                # jaxob_constants.py:9
EMPTY_SLOT = -1            # jaxob_constants.py:11
ORDERBOOK_FEAT = 6         # jaxob_constants.py:13
TRADE_FEAT = 8             # jaxob_constants.py:14
MSG_FEAT = 8               # jaxob_constants.py:76-83


class CancelMode:          # jaxob_constants.py:61-65
    STRICT_BY_ID = 0
    INCLUDE_INITS = 1
    CANCEL_UNIFORM = 2             # needs jax.random -> unsupported (VITMARL_EUNSUPPORTED)
    CANCEL_UNIFORM_AND_LARGE = 3   # idem


class SimulatorMode:       # jaxob_constants.py:71-73
    GENERAL_EXCHANGE = 0
    LOBSTER_INTERPRETER = 1        # unfinished in the reference -> unsupported


@dataclass(frozen=True)
class JAXLOB_Configuration:        # jaxob_config.py:11-29
    maxint: int = MAXINT
    init_id: int = INITID
    cancel_mode: int = CancelMode.INCLUDE_INITS
    seed: int = 42
    nTrades: int = 100
    nOrders: int = 100
    simulator_mode: int = SimulatorMode.GENERAL_EXCHANGE
    empty_slot_val: int = EMPTY_SLOT


@dataclass(frozen=True)
class World_EnvironmentConfig(JAXLOB_Configuration):   # jaxob_config.py:158-183
    n_data_msg_per_step: int = 1
    nOrdersPerSide: int = 100
    nTradesLogged: int = 100
    book_depth: int = 10
    tick_size: int = 100
    order_id_counter_start_when_resetting: int = -200
