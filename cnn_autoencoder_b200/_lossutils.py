"""Rate-distortion criterion of the hot path, mirroring the reference's
``models.criteria`` for criterion ``'RateMSE'`` (``_critargs.py:42`` default):
``RateLoss`` (``/root/reference/src/models/criteria/_ratedist.py:45-54``),
``DistMSELoss`` (:57-63), ``GeneralLoss`` / ``setup_loss``
(``_lossutils.py:5-151``).  MS-SSIM / pyramid / penalty / classification terms of
the reference are outside the compress/decompress path
and raise ``NotImplementedError`` when requested.
"""
import torch
import torch.nn as nn


class RateLoss(object):
    def __init__(self, **kwargs):
        pass

    def __call__(self, x, p_y, **kwargs):
        # normalised by the pixels of the IMAGE, not of the latent (:51-52)
        rate_loss = -torch.sum(torch.log2(p_y)) / (x.size(0) * x.size(2) * x.size(3))
        return dict(rate_loss=rate_loss)


class DistMSELoss(object):
    def __init__(self, **kwargs):
        self._mse = nn.MSELoss()

    def __call__(self, x, x_r, **kwargs):
        return dict(dist=[self._mse(x_r[0], x.to(x_r[0].device))])


DIST_LOSS_LIST = {'MSE': DistMSELoss}
RATE_LOSS_LIST = {'Rate': RateLoss}


class GeneralLoss(nn.Module):
    def __init__(self, dist_loss_type='MSE', rate_loss_type='Rate', penalty_loss_type=None,
                 class_loss_type=None, distortion_lambda=0.1, penalty_beta=0.001,
                 class_error_mu=1.0, class_error_aux_mu=1.0, **kwargs):
        super().__init__()
        if penalty_loss_type is not None and str(penalty_loss_type).lower() != 'none':
            raise NotImplementedError('penalty criteria are outside the hot path')
        if class_loss_type is not None and str(class_loss_type).lower() != 'none':
            raise NotImplementedError('classification criteria are outside the hot path')
        self.dist_loss = None
        if dist_loss_type is not None:
            if dist_loss_type not in DIST_LOSS_LIST:
                raise NotImplementedError(f'distortion {dist_loss_type!r} is outside the hot path')
            self.dist_loss = DIST_LOSS_LIST[dist_loss_type](**kwargs)
            self._multiplier = 255 ** 2 if 'MSE' in dist_loss_type else 1
            if not isinstance(distortion_lambda, list):
                distortion_lambda = [distortion_lambda]
            self._distortion_lambda = distortion_lambda
        self.rate_loss = None
        if rate_loss_type is not None:
            self.rate_loss = RATE_LOSS_LIST[rate_loss_type](**kwargs)

    def forward(self, inputs, outputs, targets=None, net=None, **kwargs):
        loss_dict = {'loss': 0, 'channel_e': torch.LongTensor([-1])}
        if self.dist_loss is not None:
            loss_dict.update(self.dist_loss(x=inputs, x_r=outputs['x_r'], **kwargs))
            loss_dict['dist'] = [self._multiplier * d for d in loss_dict['dist']]
            loss_dict['dist_loss'] = sum(d * lam for d, lam in
                                         zip(loss_dict['dist'], self._distortion_lambda))
            loss_dict['loss'] = loss_dict['loss'] + loss_dict['dist_loss']
        if self.rate_loss is not None:
            loss_dict.update(self.rate_loss(x=inputs, p_y=outputs['p_y'], **kwargs))
            loss_dict['entropy_loss'] = net['fact_ent'].module.loss()
            loss_dict['loss'] = loss_dict['loss'] + loss_dict['rate_loss']
        return loss_dict


def setup_loss(criterion, **kwargs):
    c = criterion.lower()
    rate = 'Rate' if 'rate' in c else None
    if 'mse' in c:
        dist = 'MSE'
    elif 'msssim' in c or 'ms-ssim' in c:
        dist = 'MSSSIM'
    else:
        dist = None
    if 'multiscale' in c:
        dist = 'Multiscale' + str(dist)
    if 'penaltya' in c or 'pa' in c:
        penalty = 'PenaltyA'
    elif 'penaltyb' in c or 'pb' in c:
        penalty = 'PenaltyB'
    else:
        penalty = 'none'
    if 'bce' in c or 'binarycrossentropy' in c:
        cls = 'BCELoss'
    elif 'ce' in c or 'crossentropy' in c:
        cls = 'CELoss'
    else:
        cls = None
    return GeneralLoss(dist, rate, penalty, cls, **kwargs)
