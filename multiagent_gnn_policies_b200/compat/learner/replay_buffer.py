"""Ring buffer of transitions (reference: learner/replay_buffer.py:4-49); host-side, training only."""
import random
from collections import namedtuple

Transition = namedtuple('Transition', ('state', 'action', 'done', 'next_state', 'reward'))


class ReplayBuffer(object):
    def __init__(self, max_size=1000):
        self.max_size = max_size
        self.clear()

    def insert(self, sample):
        item = Transition(*sample)
        if self.curr_size < self.max_size:
            self.buffer.append(item)
            self.curr_size += 1
        else:
            self.buffer[self.position] = item
        self.position = (self.position + 1) % self.max_size

    def sample(self, num_samples):
        return random.sample(self.buffer, num_samples)

    def clear(self):
        self.buffer = []
        self.curr_size = 0
        self.position = 0
