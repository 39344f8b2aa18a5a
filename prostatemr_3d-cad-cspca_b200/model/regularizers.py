"""tf.keras.regularizers.l2 (R:networks.py:47-48): penalty l2 * sum(w^2)."""


class L2:
    def __init__(self, l2=0.01):
        self.l2 = float(l2)

    def get_config(self):
        return {"l2": self.l2}


def l2(l2=0.01):  # noqa: A001 - keeps the Keras spelling
    return L2(l2)


def coefficient(reg):
    """l2 coefficient of our L2, a tf.keras.regularizers.L2 (duck-typed) or None."""
    if reg is None:
        return 0.0
    return float(getattr(reg, "l2"))
