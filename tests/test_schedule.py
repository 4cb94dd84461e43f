"""MCMCSchedule: the reference's only golden vector (test/runtests.jl:5-32), verbatim."""
import extensiblemcmc_jl_b200 as em

EXPECTED = (
    (1, 1), (1, 2), (1, 3), (1, 4),
    (2, 1), (2, 2), (2, 3), (2, 4),
    (3, 2), (3, 3), (3, 4),
    (4, 3), (4, 4),
    (5, 2), (5, 3), (5, 4), (5, 5), (5, 6),
    (6, 3), (6, 5), (6, 6),
    (7, 2), (7, 3), (7, 5), (7, 6),
    (8, 3), (8, 6),
    (9, 1), (9, 2), (9, 3), (9, 6),
    (10, 1), (10, 3), (10, 5), (10, 6),
)


def test_schedule_golden_sequence():
    schedule = em.MCMCSchedule(10, 4, [(1, range(3, 9)), (2, range(4, 11, 2))])
    seen = []
    for i, s in enumerate(schedule):
        assert EXPECTED[i] == (s.mcmciter, s.pidx)
        seen.append(s)
        if s.mcmciter == 5 and s.pidx == 3:
            em.reschedule_(schedule, 2, [4], [(5, range(8, 10))])
    assert len(seen) == len(EXPECTED) == 35


def test_prev_fields_follow_the_executed_steps():
    steps = list(em.MCMCSchedule(4, 3, [(2, range(2, 4))]))
    assert steps[0].prev_mcmciter is None and steps[0].prev_pidx is None
    for a, b in zip(steps, steps[1:]):
        assert (b.prev_mcmciter, b.prev_pidx) == (a.mcmciter, a.pidx)
    assert [(s.mcmciter, s.pidx) for s in steps] == [
        (1, 1), (1, 2), (1, 3), (2, 1), (2, 3), (3, 1), (3, 3), (4, 1), (4, 2), (4, 3)]


def test_initial_state_is_not_checked_for_exclusion():
    # schedule.jl:28,57 -- (1, 1) is yielded even when update 1 is excluded at iteration 1
    steps = list(em.MCMCSchedule(2, 2, [(1, range(1, 3))]))
    assert [(s.mcmciter, s.pidx) for s in steps] == [(1, 1), (1, 2), (2, 2)]


def test_empty_and_single():
    assert list(em.MCMCSchedule(0, 3)) == []
    assert [(s.mcmciter, s.pidx) for s in em.MCMCSchedule(3, 1)] == [(1, 1), (2, 1), (3, 1)]
