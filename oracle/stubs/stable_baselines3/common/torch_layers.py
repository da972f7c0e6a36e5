from torch import nn


class BaseFeaturesExtractor(nn.Module):
    def __init__(self, observation_space, features_dim=0):
        super().__init__()
        self._observation_space = observation_space
        self._features_dim = features_dim

    @property
    def features_dim(self):
        return self._features_dim


class FlattenExtractor(BaseFeaturesExtractor):
    pass
