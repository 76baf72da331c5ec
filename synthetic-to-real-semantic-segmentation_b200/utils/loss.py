"""Losses with the reference's class and method names (utils/loss.py:5-69)."""
import torch

from ..functional import cross_entropy, bce_with_logits


class SegmentationLosses(object):
    def __init__(self, weight=None, batch_average=True, ignore_index=255, cuda=False):
        self.ignore_index = ignore_index
        self.weight = weight
        self.batch_average = batch_average
        self.cuda = cuda

    def build_loss(self, mode='ce'):
        """Choices: ['ce' or 'focal']"""
        losses = {'ce': self.CrossEntropyLoss, 'focal': self.FocalLoss}
        if mode not in losses:
            raise NotImplementedError
        return losses[mode]

    def _mean_ce(self, logit, target):
        # nn.CrossEntropyLoss(weight, ignore_index, reduction='mean')(logit, target.long()), loss.py:21-30
        return cross_entropy(logit, target, weight=self.weight, ignore_index=self.ignore_index)

    def CrossEntropyLoss(self, logit, target):
        return self._mean_ce(logit, target)

    def FocalLoss(self, logit, target, gamma=2, alpha=0.5):
        # loss.py:32-46: a scalar transform of the MEAN cross entropy, -(1 - e^-CE)^gamma * alpha * CE
        log_pt = -self._mean_ce(logit, target)
        modulation = (1 - torch.exp(log_pt)) ** gamma
        if alpha is not None:
            log_pt = log_pt * alpha
        return -modulation * log_pt


class DomainLosses(object):
    def __init__(self, batch_average=True, cuda=False):
        self.batch_average = batch_average
        self.cuda = cuda
        self.device_acc = False   # True: the accuracy stays a device scalar (the reference returns .item(), loss.py:67)

    def build_loss(self):
        return self.DomainClassiferLoss

    def DomainClassiferLoss(self, src_logit, tgt_logit):
        # loss.py:57-69: CE(src, 0) + CE(tgt, 1) and the domain accuracy as a python float
        assert src_logit.size() == tgt_logit.size()
        n, c, h, w = src_logit.size()
        stats = []
        loss = cross_entropy(src_logit, None, const_target=0, ignore_index=-100, stats_out=stats) + \
            cross_entropy(tgt_logit, None, const_target=1, ignore_index=-100, stats_out=stats)
        hits = stats[0][2] + stats[1][2]
        acc = (hits / 2 / n / h / w).float()
        if self.device_acc:      # CUDA-graph capture (steps.FeatureStep.capture): no host synchronisation
            return loss, acc
        return loss, acc.item()


class BCEWithLogitsLoss(object):
    """torch.nn.BCEWithLogitsLoss() as instantiated at train_adapt.py:75."""

    def __call__(self, input, target):
        return bce_with_logits(input, target)
