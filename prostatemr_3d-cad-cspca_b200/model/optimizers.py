"""Keras optimizer / schedule descriptors consumed by M1.compile (train_model.py:113-120,
README.md:53-61). The update itself is the fused K9 kernel (m1_adam_amsgrad)."""
import math


class CosineDecayRestarts:
    """tf.keras.optimizers.schedules.CosineDecayRestarts."""

    def __init__(self, initial_learning_rate, first_decay_steps, t_mul=2.0, m_mul=1.0, alpha=0.0):
        self.initial_learning_rate = initial_learning_rate
        self.first_decay_steps = first_decay_steps
        self.t_mul, self.m_mul, self.alpha = t_mul, m_mul, alpha

    def __call__(self, step):
        completed = step / self.first_decay_steps
        if self.t_mul == 1.0:
            i_restart = math.floor(completed)
            completed -= i_restart
        else:
            i_restart = math.floor(math.log(1.0 - completed * (1.0 - self.t_mul)) / math.log(self.t_mul))
            sum_r = (1.0 - self.t_mul ** i_restart) / (1.0 - self.t_mul)
            completed = (completed - sum_r) / self.t_mul ** i_restart
        m_fac = self.m_mul ** i_restart
        cosine = 0.5 * m_fac * (1.0 + math.cos(math.pi * completed))
        return self.initial_learning_rate * ((1 - self.alpha) * cosine + self.alpha)


class Adam:
    """tf.keras.optimizers.Adam(learning_rate, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad)."""

    def __init__(self, learning_rate=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False):
        self.learning_rate = learning_rate
        self.beta_1, self.beta_2, self.epsilon, self.amsgrad = beta_1, beta_2, epsilon, amsgrad
        self.iterations = 0

    def lr(self, step):
        lr = self.learning_rate
        return float(lr(step)) if callable(lr) else float(lr)

    def lr_t(self):
        """bias-corrected step size of the NEXT update (iterations counts completed updates)."""
        t = self.iterations + 1
        return self.lr(self.iterations) * math.sqrt(1 - self.beta_2 ** t) / (1 - self.beta_1 ** t)
