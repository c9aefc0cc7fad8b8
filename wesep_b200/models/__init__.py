"""Model registry — reference wesep/models/__init__.py:10-27 (prefix dispatch by name)."""
import wesep_b200.models.bsrnn as bsrnn
import wesep_b200.models.convtasnet as convtasnet
import wesep_b200.models.dpccn as dpccn
import wesep_b200.models.tfgridnet as tfgridnet


def get_model(model_name: str):
    if model_name.startswith("ConvTasNet"):
        return getattr(convtasnet, model_name)
    if model_name == "BSRNN":
        return bsrnn.BSRNN
    if model_name == "BSRNN_Multi":
        return bsrnn.BSRNN_Multi
    if model_name == "DPCCN":
        return dpccn.DPCCN
    if model_name == "TFGridNet":
        return tfgridnet.TFGridNet
    for prefix in ("BSRNN_Feats", "BSRNN_Multi", "BSRNN", "DPCCN", "TFGridNet", "CMGAN"):
        if model_name.startswith(prefix):
            raise NotImplementedError(model_name + " is not built yet in wesep_b200 (ConvTasNet / Spex+, BSRNN, BSRNN_Multi, DPCCN and TFGridNet are)")
    print(model_name + " not found !!!")
    exit(1)
