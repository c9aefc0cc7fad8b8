from wesep_b200.modules.common.norm import ChannelWiseLayerNorm  # noqa
from wesep_b200.modules.common.norm import GlobalChannelLayerNorm  # noqa
from wesep_b200.modules.common.norm import select_norm  # noqa
from wesep_b200.modules.common.norm import FiLM  # noqa
