"""ExponentialDecrease — reference wesep/utils/schedulers.py:99-222 (host scalar math, one float per step)."""
import math


class BaseClass:

    def __init__(self, optimizer, num_epochs, epoch_iter, initial_lr, final_lr, warm_up_epoch=6, scale_ratio=1.0,
                 warm_from_zero=False):
        self.optimizer = optimizer
        self.max_iter = num_epochs * epoch_iter
        self.initial_lr = initial_lr
        self.final_lr = final_lr
        self.scale_ratio = scale_ratio
        self.current_iter = 0
        self.warm_up_iter = warm_up_epoch * epoch_iter
        self.warm_from_zero = warm_from_zero

    def get_multi_process_coeff(self):
        lr_coeff = 1.0 * self.scale_ratio
        if self.current_iter < self.warm_up_iter:
            if self.warm_from_zero:
                lr_coeff = self.scale_ratio * self.current_iter / self.warm_up_iter
            elif self.scale_ratio > 1:
                lr_coeff = (self.scale_ratio - 1) * self.current_iter / self.warm_up_iter + 1.0
        return lr_coeff

    def get_current_lr(self):
        return 0.0

    def get_lr(self):
        return self.optimizer.param_groups[0]["lr"]

    def set_lr(self):
        current_lr = self.get_current_lr()
        for param_group in self.optimizer.param_groups:
            param_group["lr"] = current_lr

    def step(self, current_iter=None):
        if current_iter is not None:
            self.current_iter = current_iter
        self.set_lr()
        self.current_iter += 1

    def state_dict(self):
        return {key: value for key, value in self.__dict__.items() if key != "optimizer"}

    def load_state_dict(self, state_dict):
        self.__dict__.update(state_dict)


class ExponentialDecrease(BaseClass):

    def get_current_lr(self):
        lr_coeff = self.get_multi_process_coeff()
        return lr_coeff * self.initial_lr * math.exp(
            (self.current_iter / self.max_iter) * math.log(self.final_lr / self.initial_lr))
