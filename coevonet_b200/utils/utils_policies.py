"""Debug opponents of the reference (``utils/utils_policies.py``).  They are only
used by the Atari *test* mode upstream and cannot sit in a device rollout."""
from random import randint


class RandomPolicy(object):
    def __init__(self, number_actions):
        self.number_actions = number_actions

    def determine_action(self, input, args):
        return randint(0, self.number_actions - 1)


class PeriodicPolicy(object):
    def __init__(self, number_actions):
        self.number_actions = number_actions
        self.i = -1

    def determine_action(self, input, args):
        self.i = (self.i + 1) % self.number_actions
        return self.i


class AlwaysFirePolicy(object):
    def determine_action(self, input, args):
        return 1
