"""Learning-rate policy with the reference's call signature (utils/lr_scheduler.py:43-70):
scheduler(optimizer, i, epoch, best_pred) writes lr into param group 0 and 10*lr into the others
(which also overwrites the Adam lr of the discriminator, train_adapt.py:133)."""
import math


class LR_Scheduler(object):
    def __init__(self, mode, base_lr, num_epochs, iters_per_epoch=0, lr_step=0, warmup_epochs=0, quiet=True):
        self.mode = mode
        self.lr = base_lr
        if mode == 'step':
            assert lr_step
        self.lr_step = lr_step
        self.iters_per_epoch = iters_per_epoch
        self.N = num_epochs * iters_per_epoch
        self.epoch = -1
        self.warmup_iters = warmup_epochs * iters_per_epoch
        self.quiet = quiet

    def lr_at(self, i, epoch):
        T = epoch * self.iters_per_epoch + i
        if self.mode == 'cos':
            lr = 0.5 * self.lr * (1 + math.cos(1.0 * T / self.N * math.pi))
        elif self.mode == 'poly':
            lr = self.lr * pow((1 - 1.0 * T / self.N), 0.9)
        elif self.mode == 'step':
            lr = self.lr * (0.1 ** (epoch // self.lr_step))
        else:
            raise NotImplementedError
        if self.warmup_iters > 0 and T < self.warmup_iters:
            lr = lr * 1.0 * T / self.warmup_iters
        assert lr >= 0
        return lr

    def __call__(self, optimizer, i, epoch, best_pred=0.0):
        lr = self.lr_at(i, epoch)
        if epoch > self.epoch:
            if not self.quiet:
                print('\n=>Epoches %i, learning rate = %.4f, previous best = %.4f' % (epoch, lr, best_pred))
            self.epoch = epoch
        groups = optimizer.param_groups
        groups[0]['lr'] = lr
        for g in groups[1:]:
            g['lr'] = lr * 10
