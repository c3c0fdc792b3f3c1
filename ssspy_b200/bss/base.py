"""Iteration driver (host mirror of ssspy/bss/base.py:10-89)."""


class IterativeMethodBase:
    """``__call__`` runs ``update_once`` ``n_iter`` times, recording the loss and calling the
    callbacks before the loop (``initial_call``) and after every iteration, exactly as
    ssspy/bss/base.py:48-77 does.  ``update_once`` stays an overridable per-iteration Python entry
    point (the reference's notebooks subclass separators and override it)."""

    def __init__(self, callbacks=None, record_loss=True):
        if callbacks is not None and callable(callbacks):
            callbacks = [callbacks]
        self.callbacks = callbacks
        self.record_loss = record_loss
        self.loss = [] if record_loss else None

    def __call__(self, *args, n_iter=100, initial_call=True, **kwargs):
        if initial_call:
            self._after_step()
        for _ in range(n_iter):
            self.update_once()
            self._after_step()

    def _after_step(self):
        if self.record_loss:
            self.loss.append(self.compute_loss())
        if self.callbacks is not None:
            for callback in self.callbacks:
                callback(self)

    def update_once(self):
        raise NotImplementedError("Implement 'update_once' method.")

    def compute_loss(self):
        raise NotImplementedError("Implement 'compute_loss' method.")
