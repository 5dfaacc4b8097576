"""Per-step forward closure, mirroring ``models.tasks._taskutils`` of the reference
(``/root/reference/src/models/tasks/_taskutils.py:46-110``): same factory name,
arguments and returned dict keys.  Only the compress/decompress modules
(encoder, fact_ent, decoder) are in scope; the classifier / segmenter heads of
the reference are accepted in ``enabled_modules`` and driven the same way if the
caller supplies them in the model dict.
"""
import torch


def _stepper(enabled, trainable, passthrough):
    """Returns fn(model, key, *args) running model[key] with grad on/off globally,
    like ``trainable_module`` / ``fixed_module`` (:5-22), or a pass-through."""
    if not enabled:
        return passthrough

    def run(model, key, *args, **kwargs):
        torch.set_grad_enabled(bool(trainable))
        try:
            return model[key](*args, **kwargs)
        finally:
            torch.set_grad_enabled(True)

    return run


def decorate_trainable_modules(trainable_modules=None, enabled_modules=None):
    if enabled_modules is None:
        enabled_modules = ['encoder', 'decoder', 'fact_ent', 'class_model', 'seg_model']
    if trainable_modules is None:
        trainable_modules = []

    def has(k):
        return k in enabled_modules

    def trains(k):
        return k in trainable_modules

    enc = _stepper(has('encoder'), trains('encoder'), lambda m, k, x, **kw: x)
    ent = _stepper(has('fact_ent'), trains('fact_ent'), lambda m, k, x, **kw: (x, None))
    dec = _stepper(has('decoder'), trains('decoder'), lambda m, k, x, **kw: (x, None))
    cls = _stepper(has('class_model'), trains('class_model'), lambda m, k, x, **kw: (None, None))
    seg = _stepper(has('seg_model'), trains('seg_model'), lambda m, k, x, **kw: (None, None))

    def forward_func(x, model):
        y = enc(model, 'encoder', x)
        y_q, p_y = ent(model, 'fact_ent', y)
        x_r, fx_brg = dec(model, 'decoder', y_q)
        t_pred, t_aux_pred = cls(model, 'class_model', y_q)
        s_pred, s_aux_pred = seg(model, 'seg_model', y_q, fx_brg=fx_brg)
        return dict(x_r=x_r, fx_brg=fx_brg, y=y, y_q=y_q, p_y=p_y, t_pred=t_pred,
                    t_aux_pred=t_aux_pred, s_pred=s_pred, s_aux_pred=s_aux_pred)

    return forward_func
