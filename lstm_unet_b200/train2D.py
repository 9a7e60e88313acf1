"""Mirror of the hot-path part of the reference's ``train2D.py`` (train2D.py:33-118,145-161,192-220): same call
order -- providers, model construction with pad_image=False, Adam, train_step, reset_states_per_batch, validation
with swapped recurrent states, and the per-step SEG measure / accuracy metrics (train2D.py:97-102,111-116) computed on the
device, and the final export for inference (train2D.py:232-240: ``model.ckpt`` as a TF2 tensor bundle +
``model_params.pickle``, what ``Inference2D.inference`` loads), and the training checkpoints (train2D.py:62-85,222-226:
restore on ``load_checkpoint``, a checkpoint every ``save_checkpoint_iteration`` steps; checkpoint.py).  TensorBoard and
AWS polling are out of scope.  Usage, as in the reference: set the module global ``params`` and call ``train()``."""
import os
import pickle

from . import Networks as Nets
from . import losses

params = None


def log_print(*args):
    print(*args)


def _is_writer_rank():
    try:
        import torch.distributed as dist
        return not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    except Exception:
        return True


def train(num_iterations=None, allreduce=None, log=log_print):
    """Returns the list of per-step training losses (floats)."""
    train_data_provider = params.train_data_provider
    val_data_provider = params.val_data_provider
    train_data_provider.start_queues(None)
    val_data_provider.start_queues(None)

    model = params.net_model(params.net_kernel_params, params.data_format, False,
                             precision=getattr(params, 'precision', 'bf16'), train=True)
    ce_loss = losses.WeightedCELoss(params.channel_axis + 1, params.class_weights)
    optimizer = Nets.Adam(lr=params.learning_rate)
    seg_measure = losses.seg_measure(params.channel_axis + 1, three_d=False)     # train2D.py:51
    metrics = {'train': {'SEG': [], 'accuracy': []}, 'val': {'SEG': [], 'accuracy': []}}
    # train2D.py:62-85: tf.train.Checkpoint(step, optimizer, net=model) + CheckpointManager, restore on request
    from . import checkpoint as ck
    ckpt = ck.Checkpoint(model, optimizer)
    if getattr(params, 'load_checkpoint', False):
        path = os.path.expanduser(params.load_checkpoint_path)
        latest = ck.latest_checkpoint(path) if os.path.isdir(path) else path
        if latest:
            ckpt.restore(latest)
            log('Restored from {}'.format(latest))
        else:
            log('Initializing from scratch.')
    manager = ck.CheckpointManager(ckpt, os.path.join(os.path.expanduser(params.experiment_save_dir), 'tf_ckpts'),
                                   max_to_keep=getattr(params, 'save_checkpoint_max_to_keep', 5))
    step = ckpt.step

    def train_step(image, label):
        softmax, predictions, loss = model.train_step(image, label, params.class_weights, optimizer, allreduce)
        metrics['train']['SEG'].append(seg_measure(label, predictions))          # train2D.py:97-102
        metrics['train']['accuracy'].append(seg_measure.last_accuracy)
        return softmax, predictions, loss

    def val_step(image, label):
        predictions, softmax = model(image, False)
        t_loss = ce_loss(label, predictions)
        metrics['val']['SEG'].append(seg_measure(label, predictions))            # train2D.py:111-116
        metrics['val']['accuracy'].append(seg_measure.last_accuracy)
        return softmax, predictions, t_loss

    losses_seen = []
    val_states = model.get_states()
    n_iter = params.num_iterations if num_iterations is None else num_iterations
    writer = _is_writer_rank()                    # data-parallel runs: one rank writes checkpoints / the inference model
    dry = getattr(params, 'dry_run', True)
    train.model = model
    train.metrics = metrics
    try:
        for _ in range(step, n_iter):
            image_sequence, seg_sequence, _, is_last_batch = train_data_provider.get_batch()
            _, _, train_loss_value = train_step(image_sequence, seg_sequence)
            step += 1
            ckpt.step = step
            model.reset_states_per_batch(is_last_batch)      # reset states for sequences that ended (train2D.py:161)
            losses_seen.append(float(train_loss_value))
            if not step % params.print_to_console_interval:
                log('Training: Step {}, Loss: {}'.format(step, losses_seen[-1]))
            # train2D.py:222-226: every save_checkpoint_iteration steps AND at the last step
            if not dry and writer and (not step % getattr(params, 'save_checkpoint_iteration', 5000) or step == n_iter):
                log('Saved checkpoint for step {}: {}'.format(step, manager.save(step)))
            if not step % params.validation_interval:
                train_states = model.get_states()
                model.set_states(val_states)
                val_image_sequence, val_seg_sequence, _, val_is_last_batch = val_data_provider.get_batch()
                _, _, val_loss_value = val_step(val_image_sequence, val_seg_sequence)
                model.reset_states_per_batch(val_is_last_batch)
                log('Validation: Step {}, Loss: {}'.format(step, float(val_loss_value)))
                val_states = model.get_states()
                model.set_states(train_states)
    except (KeyboardInterrupt, ValueError) as err:           # train2D.py:225-230: checkpoint, then leave the loop quietly
        if not dry and writer:
            log('Saving Model Before closing due to error: {}'.format(str(err)))
            log('Saved checkpoint for step {}: {}'.format(step, manager.save(step)))
    finally:                                                 # train2D.py:232-240: the inference model, whatever happened
        if not dry and writer and model._sess is not None:
            save_dir = os.path.expanduser(params.experiment_save_dir)
            os.makedirs(save_dir, exist_ok=True)
            model_fname = os.path.join(save_dir, 'model.ckpt')
            model.save_weights(model_fname, save_format='tf')
            with open(os.path.join(save_dir, 'model_params.pickle'), 'wb') as fobj:
                pickle.dump({'name': model.__class__.__name__, 'params': (params.net_kernel_params,)}, fobj,
                            protocol=pickle.HIGHEST_PROTOCOL)
            log('Saved Model to file: {}'.format(model_fname))
        elif dry:
            log('WARNING: dry_run flag is ON! Not Saving Model')
    return losses_seen
