from torch import nn


class PreTrainedPolicy(nn.Module):
    config_class = None
    name = None

    def __init__(self, config, *inputs, **kwargs):
        super().__init__()
        self.config = config
