"""Host-side model of the activation-stage ring of lu_conv_tc_kernel (csrc/lu_conv.cuh): one in-order producer, mbarrier
parity waits (a wait names only the PARITY of the phase it wants), and one or two MMA-issuing consumers.

Why it exists: the experimental second issuing thread (LU_TWO_ISSUERS) hung on hardware in the per-tap `direct`
staging mode (27 stages per tile, 8 ring slots).  Two consumers that skip each other's slots of ONE ring without looking at
their barriers are sound only while the ring is DEEPER than one tile's stages: otherwise a consumer can reach a slot two
laps after the producer last filled it and the parity wait aliases -- it passes before the data of that lap has landed.
The model replays the protocol under random schedules and pins the rule the guard in `launch_conv` (lu_api.cu) applies:
two issuers only when n_a_stages > n_astages."""
import random

import pytest


class Bar:
    """mbarrier with arrival count 1: `done` completed phases; wait(parity) succeeds iff the phase of that parity is the
    last completed one or the barrier is one phase further (hardware: current phase parity != waited parity)."""

    def __init__(self):
        self.done = 0

    def ready(self, parity):
        return (self.done & 1) != parity


def simulate(n_slots, per_tile, tiles, issuers, seed, max_steps=200000):
    """-> (outcome, detail): 'ok', 'alias' (a consumer read a slot whose lap had not been produced) or 'deadlock'."""
    rng = random.Random(seed)
    full = [Bar() for _ in range(n_slots)]
    empty = [Bar() for _ in range(n_slots)]
    produced = {}                                     # (slot, lap) -> True once the producer filled it
    prod_pos = 0
    total = tiles * per_tile
    cons = [{'tile': w, 'stage': 0} for w in range(issuers)]
    for _ in range(max_steps):
        moved = False
        agents = ['p'] + list(range(issuers))
        rng.shuffle(agents)
        for a in agents:
            if a == 'p':
                if prod_pos >= total:
                    continue
                slot, lap = prod_pos % n_slots, prod_pos // n_slots
                if empty[slot].ready((lap & 1) ^ 1):              # the consumer of the previous lap released the slot
                    produced[(slot, lap)] = True
                    full[slot].done += 1
                    prod_pos += 1
                    moved = True
            else:
                c = cons[a]
                if c['tile'] >= tiles:
                    continue
                pos = c['tile'] * per_tile + c['stage']
                slot, lap = pos % n_slots, pos // n_slots
                if full[slot].ready(lap & 1):
                    if not produced.get((slot, lap)):
                        return 'alias', (a, c['tile'], c['stage'], slot, lap)
                    empty[slot].done += 1                       # tcgen05.commit of this stage's MMAs
                    c['stage'] += 1
                    if c['stage'] == per_tile:
                        c['stage'] = 0
                        c['tile'] += issuers                    # the other issuers' tiles are skipped without a wait
                    moved = True
        if prod_pos >= total and all(c['tile'] >= tiles for c in cons):
            return 'ok', None
        if not moved:
            return 'deadlock', (prod_pos, [dict(c) for c in cons])
    return 'timeout', None


@pytest.mark.parametrize('n_slots,per_tile', [(2, 1), (3, 1), (8, 1), (8, 3), (6, 4), (3, 9), (8, 27)])
def test_single_issuer_ring_is_sound_for_any_depth(n_slots, per_tile):
    for seed in range(20):
        assert simulate(n_slots, per_tile, 23, 1, seed)[0] == 'ok'


@pytest.mark.parametrize('n_slots,per_tile', [(2, 1), (3, 1), (5, 1), (8, 1), (4, 2), (8, 3), (8, 4), (6, 5), (8, 7), (12, 3)])
def test_two_issuers_are_sound_when_the_ring_is_deeper_than_a_tile(n_slots, per_tile):
    assert n_slots > per_tile                                # the guard of launch_conv
    for seed in range(50):
        assert simulate(n_slots, per_tile, 24, 2, seed)[0] == 'ok'
        assert simulate(n_slots, per_tile, 25, 2, seed)[0] == 'ok'


@pytest.mark.parametrize('n_slots,per_tile', [(8, 27), (8, 9), (3, 4), (2, 2), (1, 1), (8, 8)])
def test_two_issuers_alias_when_a_tile_fills_the_ring(n_slots, per_tile):
    assert n_slots <= per_tile
    outcomes = {simulate(n_slots, per_tile, 24, 2, seed)[0] for seed in range(200)}
    assert 'alias' in outcomes, outcomes                     # on hardware: stale operands, then a hang


def test_the_rule_is_exact_in_the_model():
    for n_slots in range(1, 10):
        for per_tile in range(1, 10):
            outcomes = set()
            for seed in range(40):
                outcomes.add(simulate(n_slots, per_tile, 24, 2, seed)[0])
                outcomes.add(simulate(n_slots, per_tile, 25, 2, seed)[0])
            assert (outcomes == {'ok'}) == (n_slots > per_tile), (n_slots, per_tile, outcomes)
