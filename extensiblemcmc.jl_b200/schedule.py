"""MCMCSchedule: the iterator over (mcmciter, pidx) update steps.

Host-side restatement of src/schedule.jl:17-118 with identical iteration semantics
(checked against the reference's golden sequence, test/runtests.jl:13-31).  In the GPU
backend it doubles as the block planner's source: `run.__run` pulls schedule elements
from it and ships them to the device in blocks (one CUDA graph launch per block).

Indices are 1-based exactly as in the reference (`mcmciter` enters the adaptation rule,
src/transition_kernels/adaptation.jl:312-319); the C ABI receives `pidx - 1`.
"""
from collections import namedtuple

Step = namedtuple("Step", ["prev_mcmciter", "prev_pidx", "mcmciter", "pidx"])


class MCMCSchedule:
    def __init__(self, num_mcmc_steps, num_updates, exclude_updates=(), start=None,
                 backend=None, extra_info=None):
        # schedule.jl:24-45; exclude_updates: [(update indices, iteration range), ...];
        # DefaultDict default 0:0 never matches an iteration (schedule.jl:32)
        self.num_mcmc_steps = int(num_mcmc_steps)
        self.num_updates = int(num_updates)
        self.start = start if start is not None else Step(None, None, 1, 1)
        self.exclude_updates = {}
        for idxs, rng in exclude_updates:
            for idx in ([idxs] if isinstance(idxs, int) else idxs):
                self.exclude_updates[idx] = rng
        self.extra_info = extra_info

    def _excluded(self, pidx, mcmciter):
        rng = self.exclude_updates.get(pidx)
        return rng is not None and mcmciter in rng

    def transition(self, state):
        # schedule.jl:77-89 (recursion unrolled)
        while True:
            reset = state.pidx == self.num_updates
            state = Step(state.prev_mcmciter, state.prev_pidx,
                         state.mcmciter + (1 if reset else 0), 1 if reset else state.pidx + 1)
            if not self._excluded(state.pidx, state.mcmciter):
                return state

    def extra_transitions(self, new_state):
        # schedule.jl:68
        return new_state

    def __iter__(self):
        # schedule.jl:56-66: the next state is computed BEFORE the current one is handed to
        # the loop body, so a reschedule!() issued while processing step s first shows at
        # the step after s + 1.  The initial state is yielded without an exclusion check.
        state = self.start
        while state.mcmciter <= self.num_mcmc_steps:
            tmp = Step(state.mcmciter, state.pidx, state.mcmciter, state.pidx)
            nxt = self.extra_transitions(self.transition(tmp))
            yield state
            state = nxt


def reschedule_(schedule, num_new_updates=0, idxes_to_remove=(), idxes_to_add=()):
    """reschedule!(schedule, ...) src/schedule.jl:105-118."""
    schedule.num_updates += num_new_updates
    for idx in idxes_to_remove:
        schedule.exclude_updates[idx] = range(1, schedule.num_mcmc_steps + 1)
    for idx, rng in idxes_to_add:
        schedule.exclude_updates[idx] = rng
