from wesep_b200.modules.tasnet.decoder import MultiDecoder  # noqa
from wesep_b200.modules.tasnet.encoder import MultiEncoder  # noqa
from wesep_b200.modules.tasnet.separation import Separation, FuseSeparation  # noqa
from wesep_b200.modules.tasnet.speaker import ResNet4SpExplus  # noqa
