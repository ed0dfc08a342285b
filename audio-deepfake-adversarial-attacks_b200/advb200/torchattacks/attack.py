"""Attack base class — the API shell of adversarial_attacks/torchattacks/attack.py:5-331."""
import torch

from .. import engine


class Attack(object):
    def __init__(self, name, model):
        self.attack = name
        self.model = model
        self.model_name = str(model).split("(")[0]
        self.device = next(model.parameters()).device

        self._attack_mode = "default"
        self._targeted = False
        self._return_type = "float"
        self._supported_mode = ["default"]

        self._model_training = False
        self._batchnorm_training = False
        self._dropout_training = False
        self._fused_minmax = False  # aa.attack_minmax: raw waveforms in / out, min-max scaling inside the native call

    def forward(self, *input):
        raise NotImplementedError

    def get_mode(self):
        return self._attack_mode

    def set_mode_default(self):
        self._attack_mode = "default"
        self._targeted = False
        print("Attack mode is changed to 'default.'")

    def set_mode_targeted_by_function(self, target_map_function=None):
        raise NotImplementedError("targeted modes are not used by the repo (SURVEY.md §8 f4) and not built natively")

    set_mode_targeted_least_likely = set_mode_targeted_random = set_mode_targeted_by_function

    def set_return_type(self, type):
        if type == "float":
            self._return_type = "float"
        elif type == "int":
            self._return_type = "int"
        else:
            raise ValueError(type + " is not a valid type. [Options: float, int]")

    def set_training_mode(self, model_training=False, batchnorm_training=False, dropout_training=False):
        """attack.py:132-147.  The native forward is the eval-mode forward (running BN stats, identity dropout),
        which is what every call site of the repo selects (``set_training_mode(True, False)``)."""
        if batchnorm_training or dropout_training:
            raise NotImplementedError("advb200 implements eval-mode BatchNorm/Dropout during attacks only")
        self._model_training = model_training
        self._batchnorm_training = batchnorm_training
        self._dropout_training = dropout_training

    def _engine(self, images):
        return engine.engine_for(self.model, images.shape[0], images.shape[1])

    def _prepare(self, images, labels):
        images = images.clone().detach().to(self.device)
        labels = labels.clone().detach().to(self.device)
        if images.device.type != "cuda":
            raise RuntimeError("advb200 has no CPU path: put the model on a CUDA device")
        return images, labels

    def _to_uint(self, images):
        return (images * 255).type(torch.uint8)

    def __str__(self):
        info = self.__dict__.copy()
        for key in [k for k in info if k[0] == "_"] + ["model", "attack"]:
            info.pop(key, None)
        info["attack_mode"] = self._attack_mode
        info["return_type"] = self._return_type
        return self.attack + "(" + ", ".join("{}={}".format(k, v) for k, v in info.items()) + ")"

    def __call__(self, *input, **kwargs):
        # attack.py:308-331: the reference flips train()/eval() flags around the call; the native engine never
        # runs a PyTorch module, so only the restore semantics matter: model.training is left as it was.
        given_training = self.model.training
        images = self.forward(*input, **kwargs)
        if given_training != self.model.training:
            self.model.train(given_training)
        if self._return_type == "int":
            images = self._to_uint(images)
        return images
