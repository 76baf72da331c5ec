"""Learning-rate policy with the reference's call signature (utils/lr_scheduler.py:43-70):
scheduler(optimizer, i, epoch, best_pred) writes lr into param group 0 and 10*lr into the others
(which also overwrites the Adam lr of the discriminator, train_adapt.py:133).

The three policies are functions of the global iteration T = epoch * iters_per_epoch + i, the total
N = num_epochs * iters_per_epoch and the epoch; each keeps the reference's order of floating-point
operations, so the values are bit-identical (tests/golden/policy.npz)."""
import math

_POLICIES = {
    # lr_scheduler.py:45-46 / :47-48 / :49-50
    'cos': lambda base, T, N, epoch, lr_step: 0.5 * base * (1 + math.cos(1.0 * T / N * math.pi)),
    'poly': lambda base, T, N, epoch, lr_step: base * pow((1 - 1.0 * T / N), 0.9),
    'step': lambda base, T, N, epoch, lr_step: base * (0.1 ** (epoch // lr_step)),
}


class LR_Scheduler(object):
    def __init__(self, mode, base_lr, num_epochs, iters_per_epoch=0, lr_step=0, warmup_epochs=0, quiet=True):
        if mode == 'step' and not lr_step:
            raise AssertionError("step mode needs lr_step")            # `assert lr_step` at lr_scheduler.py:35
        self.mode, self.lr, self.lr_step = mode, base_lr, lr_step
        self.iters_per_epoch = iters_per_epoch
        self.N = num_epochs * iters_per_epoch
        self.warmup_iters = warmup_epochs * iters_per_epoch
        self.epoch = -1                      # last epoch announced
        self.quiet = quiet                   # the reference prints a line per epoch (and one at construction)

    def lr_at(self, i, epoch):
        """The learning rate of iteration i of `epoch` (what the reference computes inside __call__)."""
        policy = _POLICIES.get(self.mode)
        if policy is None:
            raise NotImplementedError(self.mode)
        T = epoch * self.iters_per_epoch + i
        lr = policy(self.lr, T, self.N, epoch, self.lr_step)
        if self.warmup_iters > 0 and T < self.warmup_iters:        # linear warm-up, :54-55
            lr = lr * 1.0 * T / self.warmup_iters
        assert lr >= 0
        return lr

    def __call__(self, optimizer, i, epoch, best_pred=0.0):
        lr = self.lr_at(i, epoch)
        if epoch > self.epoch:
            self.epoch = epoch
            if not self.quiet:
                print('\n=>Epoches %i, learning rate = %.4f, previous best = %.4f' % (epoch, lr, best_pred))
        # :63-70: group 0 gets lr, every further group ("the head") ten times that
        for k, group in enumerate(optimizer.param_groups):
            group['lr'] = lr if k == 0 else lr * 10
