"""The slice of the tf.keras surface that the reference's training driver touches (train.py:128-172,
nb-radial cell 10): optimizer/loss objects handed to ``model.compile``, the ``Callback`` protocol and ``History``.
These are thin descriptions -- the arithmetic they name runs in the CUDA library."""
from __future__ import annotations


class Adam:
    """tf.keras.optimizers.Adam(learning_rate, beta_1, beta_2, epsilon) -- Keras defaults; epsilon is applied
    outside the bias correction (dib_adam_step)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, name="adam"):
        self.learning_rate = learning_rate
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.name = name

    lr = property(lambda self: self.learning_rate, lambda self, v: setattr(self, "learning_rate", v))


class optimizers:
    Adam = Adam

    @staticmethod
    def get(identifier):
        """tf.keras.optimizers.get('adam') as used at train.py:128."""
        if isinstance(identifier, Adam):
            return identifier
        if isinstance(identifier, str) and identifier.lower() == "adam":
            return Adam()
        raise ValueError(f"only Adam is implemented by the B200 engine, got {identifier!r}")


class _Loss:
    kind = None

    def __init__(self, from_logits=False, name=None):
        self.from_logits = from_logits
        self.name = name


class BinaryCrossentropy(_Loss):
    """data.py:65 / nb-radial: BinaryCrossentropy(from_logits=True)."""
    kind = "bce_logits"


class SparseCategoricalCrossentropy(_Loss):
    """data.py:343: SparseCategoricalCrossentropy(from_logits=True)."""
    kind = "sparse_ce_logits"


class MeanSquaredError(_Loss):
    kind = "mse"

    def __init__(self, name=None):
        super().__init__(from_logits=True, name=name)


class losses:
    BinaryCrossentropy = BinaryCrossentropy
    SparseCategoricalCrossentropy = SparseCategoricalCrossentropy
    MeanSquaredError = MeanSquaredError


def resolve_loss(loss):
    if isinstance(loss, _Loss):
        if not loss.from_logits:
            raise NotImplementedError(
                "the fused loss kernels take logits (from_logits=True), as every reference call site does")
        return loss.kind
    if isinstance(loss, str):
        key = loss.lower()
        if key in ("mse", "mean_squared_error"):
            return "mse"
        if key in ("bce_logits", "sparse_ce_logits"):
            return key
        if key in ("external", "custom"):       # GradientTape-style loops: the caller owns the task loss
            return "external"
    raise ValueError(f"unsupported loss {loss!r}")


class Callback:
    """tf.keras.callbacks.Callback protocol (models.py:125,152,188): the trainer sets ``.model``."""
    model = None

    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): pass
    def on_train_end(self, logs=None): pass
    def on_epoch_begin(self, epoch, logs=None): pass
    def on_epoch_end(self, epoch, logs=None): pass
    def on_train_batch_begin(self, batch, logs=None): pass
    def on_train_batch_end(self, batch, logs=None): pass


class callbacks:
    Callback = Callback


class History(Callback):
    """Return value of Model.fit: ``.history`` maps metric name -> list over epochs (train.py:169-172)."""

    def __init__(self):
        self.history = {}
        self.epoch = []

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)
