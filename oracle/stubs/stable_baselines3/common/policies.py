from torch import nn


class ActorCriticPolicy(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
