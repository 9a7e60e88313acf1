"""Mirror of the model-call part of the reference's ``Inference2D.inference`` (Inference2D.py:25-62): the model is
built with pad_image=True and called once per frame with B=1, T=1, the recurrence living in the stateful h/c; the
first ``pre_sequence_frames`` frames, played in reverse, warm the state up (DataHandeling.py:1583-1585).  The CPU
instance-labelling post-processing and TIFF output (Inference2D.py:64-131) are out of scope: each
frame's soft-max is handed to ``on_frame(t, softmax_numpy)`` instead.  Usage: set ``params`` and call ``inference()``."""
import os
import pickle

import numpy as np

from . import Networks as Nets

params = None


def get_model(name):
    return getattr(Nets, name)            # utils.py:38-40


def inference(frames=None, on_frame=None, model=None):
    if model is None:
        with open(os.path.join(params.model_path, 'model_params.pickle'), 'rb') as fobj:
            model_dict = pickle.load(fobj)
        model_cls = get_model(model_dict['name'])
        model = model_cls(*model_dict['params'], data_format=params.data_format, pad_image=True,
                          precision=getattr(params, 'precision', 'bf16'))
        model.load_weights(os.path.join(params.model_path, 'model.ckpt'))
    frames = list(frames if frames is not None else params.data_reader)
    pre = params.pre_sequence_frames
    sequence = frames[:pre][::-1] + frames
    outputs = []
    for T, image in enumerate(sequence):
        t = T - pre
        image = np.asarray(image, dtype=np.float32)
        if image.ndim != 2:
            raise ValueError()
        if params.data_format == 'NCHW':
            image = image.reshape(1, 1, 1, image.shape[0], image.shape[1])
        else:
            image = image.reshape(1, 1, image.shape[0], image.shape[1], 1)
        _, image_softmax = model(image, training=False)
        image_softmax_np = np.squeeze(image_softmax.numpy(), (0, 1))
        if t < 0:
            continue
        outputs.append(image_softmax_np.copy())
        if on_frame is not None:
            on_frame(t, image_softmax_np)
    return outputs
