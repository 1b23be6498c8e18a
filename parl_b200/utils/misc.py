"""check_model_method (parl/utils/utils.py:217-243) and machine probes (parl/utils/machine_info.py)."""
import torch


def check_model_method(model, method, algo):
    if method == 'forward':
        assert callable(getattr(model, 'forward', None)), "forward should be a function in model class"
        assert model.forward.__func__ is not super(model.__class__, model).forward.__func__, \
            "{}'s model needs to implement forward method. \n".format(algo)
    else:
        assert hasattr(model, method) and callable(getattr(model, method, None)), \
            "{}'s model needs to implement {} method. \n".format(algo, method)


def is_gpu_available():
    return torch.cuda.is_available()


def get_gpu_count():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0
