"""diffusers.callbacks: the two callback base classes pipeline_chronoedit.py:25,526 tests its `callback_on_step_end` against."""


class PipelineCallback:
    tensor_inputs = []

    def __init__(self, cutoff_step_ratio=1.0, cutoff_step_index=None):
        self.cutoff_step_ratio, self.cutoff_step_index = cutoff_step_ratio, cutoff_step_index

    def callback_fn(self, pipeline, step_index, timesteps, callback_kwargs):
        raise NotImplementedError

    def __call__(self, pipeline, step_index, timestep, callback_kwargs):
        return self.callback_fn(pipeline, step_index, timestep, callback_kwargs)


class MultiPipelineCallbacks:
    def __init__(self, callbacks):
        self.callbacks = callbacks

    @property
    def tensor_inputs(self):
        return [i for c in self.callbacks for i in c.tensor_inputs]

    def __call__(self, pipeline, step_index, timestep, callback_kwargs):
        for c in self.callbacks:
            callback_kwargs = c(pipeline, step_index, timestep, callback_kwargs)
        return callback_kwargs
