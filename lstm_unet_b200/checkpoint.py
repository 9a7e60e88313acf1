"""Training checkpoints without TensorFlow: mirrors of ``tf.train.Checkpoint(step, optimizer, net=model)``,
``tf.train.CheckpointManager`` and ``tf.train.latest_checkpoint`` as train2D.py uses them (train2D.py:62-85,222-226),
on top of the tensor-bundle reader / writer of tf_checkpoint.py.

Keys follow TensorFlow's object graph naming: ``net/<variable path>/.ATTRIBUTES/VARIABLE_VALUE`` for the model variables,
``net/<variable path>/.OPTIMIZER_SLOT/optimizer/{m,v}/.ATTRIBUTES/VARIABLE_VALUE`` for the Adam moments,
``optimizer/iter/...`` and ``step/...`` (int64) for the counters; the directory carries TensorFlow's ``checkpoint`` state
file (``model_checkpoint_path: "ckpt-N"``).  Host-side code: everything here runs without a GPU."""
import os
import re

import numpy as np

from . import tf_checkpoint as tfc

_SLOT = '/.OPTIMIZER_SLOT/optimizer/%s' + tfc.VAR_SUFFIX


def _var_path(name, root='net/'):
    return tfc.keras_key(name, root)[:-len(tfc.VAR_SUFFIX)]


class Checkpoint:
    def __init__(self, net, optimizer=None, step=0):
        self.net, self.optimizer, self.step = net, optimizer, int(step)

    def write(self, prefix):
        weights = self.net.get_weights_dict() if self.net._sess is not None else self.net._pending_weights
        if weights is None:
            raise RuntimeError('nothing to save: the model has no weights yet')
        tensors = {tfc.keras_key(n, 'net/'): v for n, v in weights.items()}
        tensors['step' + tfc.VAR_SUFFIX] = np.int64(self.step)
        if self.optimizer is not None:
            it, m, v = self.optimizer.get_slots()
            tensors['optimizer/iter' + tfc.VAR_SUFFIX] = np.int64(it)
            tensors['optimizer/learning_rate' + tfc.VAR_SUFFIX] = np.float32(self.optimizer.lr)
            tensors['optimizer/beta_1' + tfc.VAR_SUFFIX] = np.float32(self.optimizer.beta_1)
            tensors['optimizer/beta_2' + tfc.VAR_SUFFIX] = np.float32(self.optimizer.beta_2)
            if m is not None:
                for e in self.net._variable_layout():
                    if e['trainable']:
                        sl = slice(e['offset'], e['offset'] + e['count'])
                        tensors[_var_path(e['name']) + _SLOT % 'm'] = m[sl].reshape(e['shape'])
                        tensors[_var_path(e['name']) + _SLOT % 'v'] = v[sl].reshape(e['shape'])
        tfc.write_bundle(prefix, tensors)
        return prefix

    def restore(self, prefix):
        """Model variables always; step / optimizer state when the file has them (a ``save_weights`` file has not)."""
        tensors = tfc.read_bundle(prefix)
        layout = self.net._variable_layout()
        self.net.set_weights_dict(tfc.load_model_weights(prefix, [e['name'] for e in layout]))
        if 'step' + tfc.VAR_SUFFIX in tensors:
            self.step = int(np.asarray(tensors['step' + tfc.VAR_SUFFIX]).reshape(-1)[0])
        if self.optimizer is not None and 'optimizer/iter' + tfc.VAR_SUFFIX in tensors:
            n = sum(e['count'] for e in layout if e['trainable'])
            m, v, have = np.zeros(n, np.float32), np.zeros(n, np.float32), False
            for e in layout:
                km = _var_path(e['name']) + _SLOT % 'm'
                if e['trainable'] and km in tensors:
                    sl = slice(e['offset'], e['offset'] + e['count'])
                    m[sl] = tensors[km].reshape(-1)
                    v[sl] = tensors[_var_path(e['name']) + _SLOT % 'v'].reshape(-1)
                    have = True
            self.optimizer.set_slots(int(np.asarray(tensors['optimizer/iter' + tfc.VAR_SUFFIX]).reshape(-1)[0]), m if have else None, v if have else None)
            # tf.train.Checkpoint restores the optimizer's hyper-parameters with it
            for attr, key in (('lr', 'learning_rate'), ('beta_1', 'beta_1'), ('beta_2', 'beta_2')):
                k = 'optimizer/' + key + tfc.VAR_SUFFIX
                if k in tensors:
                    setattr(self.optimizer, attr, float(np.asarray(tensors[k]).reshape(-1)[0]))
        return self


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the directory's ``checkpoint`` state file, or None."""
    state = os.path.join(directory, 'checkpoint')
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    p = p if os.path.isabs(p) else os.path.join(directory, p)
    return p if os.path.exists(p + '.index') else None


class CheckpointManager:
    def __init__(self, checkpoint, directory, max_to_keep=5, keep_checkpoint_every_n_hours=None, checkpoint_name='ckpt'):
        self.checkpoint, self.directory, self.max_to_keep, self.name = checkpoint, directory, max_to_keep, checkpoint_name
        self.checkpoints = []
        state = os.path.join(directory, 'checkpoint')
        if os.path.exists(state):
            for p in re.findall(r'^all_model_checkpoint_paths:\s*"(.*)"', open(state).read(), re.M):
                p = p if os.path.isabs(p) else os.path.join(directory, p)
                if os.path.exists(p + '.index'):
                    self.checkpoints.append(p)

    @property
    def latest_checkpoint(self):
        return self.checkpoints[-1] if self.checkpoints else None

    def save(self, checkpoint_number=None):
        n = self.checkpoint.step if checkpoint_number is None else int(checkpoint_number)
        os.makedirs(self.directory, exist_ok=True)
        prefix = os.path.join(self.directory, '%s-%d' % (self.name, n))
        self.checkpoint.write(prefix)
        if prefix in self.checkpoints:
            self.checkpoints.remove(prefix)
        self.checkpoints.append(prefix)
        while self.max_to_keep and len(self.checkpoints) > self.max_to_keep:
            old = self.checkpoints.pop(0)
            for suffix in ('.index', '.data-00000-of-00001'):
                if os.path.exists(old + suffix):
                    os.remove(old + suffix)
        with open(os.path.join(self.directory, 'checkpoint'), 'w') as f:
            f.write('model_checkpoint_path: "%s"\n' % os.path.basename(prefix))
            for p in self.checkpoints:
                f.write('all_model_checkpoint_paths: "%s"\n' % os.path.basename(p))
        return prefix
