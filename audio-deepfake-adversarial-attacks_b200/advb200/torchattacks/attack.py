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
        """attack.py:60-77.  ``None`` uses the given labels as target labels."""
        if "targeted" not in self._supported_mode:
            raise ValueError("Targeted mode is not supported.")
        self._attack_mode = "targeted"
        self._targeted = True
        self._target_map_function = target_map_function
        print("Attack mode is changed to 'targeted.'")

    def set_mode_targeted_least_likely(self, kth_min=1):
        """attack.py:79-94."""
        if "targeted" not in self._supported_mode:
            raise ValueError("Targeted mode is not supported.")
        self._attack_mode = "targeted(least-likely)"
        self._targeted = True
        assert kth_min > 0
        self._kth_min = kth_min
        self._target_map_function = self._get_least_likely_label
        print("Attack mode is changed to 'targeted(least-likely).'")

    def set_mode_targeted_random(self):
        """attack.py:96-108."""
        if "targeted" not in self._supported_mode:
            raise ValueError("Targeted mode is not supported.")
        self._attack_mode = "targeted(random)"
        self._targeted = True
        self._target_map_function = self._get_random_target_label
        print("Attack mode is changed to 'targeted(random).'")

    @torch.no_grad()
    def _get_target_label(self, images, labels=None):
        """attack.py:258-270 (the model.eval()/train() flips are no-ops here: the engine always runs the eval forward)."""
        if not self._targeted:
            raise ValueError("Please define target_map_function.")
        if self._target_map_function is None:  # "None for using input labels as targeted labels"
            return labels
        return self._target_map_function(images, labels)

    @torch.no_grad()
    def _get_least_likely_label(self, images, labels=None):
        """attack.py:273-287 on the model's own output.  The detectors emit ONE logit (B,1), so ``outputs.shape[-1]`` is 1
        and the reference's ``list(range(1)).remove(label)`` raises for every label but 0: the
        least-likely / random target pickers never worked on these models.  Same failure here, stated plainly."""
        raise ValueError("targeted(least-likely) / targeted(random) pick targets over outputs.shape[-1] classes; the "
                         "detectors emit a single logit (B,1), where the reference's picker fails too (attack.py:280). "
                         "Use set_mode_targeted_by_function(lambda images, labels: 1 - labels).")

    _get_random_target_label = _get_least_likely_label

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
