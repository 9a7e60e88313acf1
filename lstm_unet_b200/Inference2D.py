"""Mirror of the model-call part of the reference's ``Inference2D.inference`` (Inference2D.py:25-62): the model is
built with pad_image=True and called once per frame with B=1, T=1, the recurrence living in the stateful h/c; the
first ``pre_sequence_frames`` frames, played in reverse, warm the state up (DataHandeling.py:1583-1585).  Unless
``params.dry_run``, the instance labelling that follows in the reference (Inference2D.py:64-123: numpy / SciPy / OpenCV
on the host) runs on the device from the soft-max that is still in HBM (postprocess.PostProcessor -> lu_postprocess,
bit-identical results); the uint16 label image goes to ``on_labels(t, labels)``, to ``last_labels`` and -- if
``params.output_path`` is set and OpenCV is importable -- to ``mask{t:03d}.tif`` like Inference2D.py:124-126.  Each
frame's soft-max is handed to ``on_frame(t, softmax_numpy)``.  Usage: set ``params`` and call ``inference()``."""
import os
import pickle

import numpy as np

from . import Networks as Nets

params = None
last_labels = []


def get_model(name):
    return getattr(Nets, name)            # utils.py:38-40


def inference(frames=None, on_frame=None, model=None, on_labels=None):
    global last_labels
    if model is None:
        with open(os.path.join(params.model_path, 'model_params.pickle'), 'rb') as fobj:
            model_dict = pickle.load(fobj)
        model_cls = get_model(model_dict['name'])
        model = model_cls(*model_dict['params'], data_format=params.data_format, pad_image=True,
                          precision=getattr(params, 'precision', 'bf16'))
        model.load_weights(os.path.join(params.model_path, 'model.ckpt'))
    pre = params.pre_sequence_frames
    if frames is not None:                # in-memory frames: the warm-up prefix is built here
        frames = list(frames)
        sequence = frames[:pre][::-1] + frames
    else:                                 # Inference2D.py:42-43: the reader's dataset already starts with the prefix
        sequence = params.data_reader(params.sequence_path, params.filename_format, pre_sequence_frames=pre).dataset
    outputs = []
    last_labels = []
    post = None
    if not getattr(params, 'dry_run', False):
        from .postprocess import PostProcessor
        post = PostProcessor(params)
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
        if post is not None:
            labels_out = post(image_softmax[0, 0]).numpy().copy()
            last_labels.append(labels_out)
            if on_labels is not None:
                on_labels(t, labels_out)
            out_dir = getattr(params, 'output_path', None)
            if out_dir:
                try:
                    import cv2
                except ImportError:
                    cv2 = None
                if cv2 is not None:
                    os.makedirs(out_dir, exist_ok=True)
                    cv2.imwrite(os.path.join(out_dir, 'mask{time:03d}.tif'.format(time=t)), labels_out)
                    if getattr(params, 'save_intermediate', False):          # Inference2D.py:105-112,127-130
                        vis_dir = getattr(params, 'save_intermediate_vis_path', os.path.join(out_dir, 'Softmax'))
                        lab_dir = getattr(params, 'save_intermediate_label_path', os.path.join(out_dir, 'Labels'))
                        os.makedirs(vis_dir, exist_ok=True)
                        os.makedirs(lab_dir, exist_ok=True)
                        sm_hwc = image_softmax_np if params.data_format != 'NCHW' else np.transpose(image_softmax_np, (1, 2, 0))
                        vis = np.flip(np.round(sm_hwc * (2 ** 16 - 1)).astype(np.uint16), 2)
                        cv2.imwrite(os.path.join(vis_dir, 'softmax{time:03d}.tif'.format(time=t)), vis)
                        cv2.imwrite(os.path.join(lab_dir, 'mask{time:03d}.tif'.format(time=t)), labels_out)
    return outputs
