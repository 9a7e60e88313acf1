"""Mirror of the reference's ``Params.py`` config surface (Params.py:14-195): same class names, attribute names and
``params_dict`` override semantics.  Differences, all at the edge of the hot path: data providers default to the
synthetic provider and are constructed lazily (the reference instantiates real CTC readers in ``__init__``,
Params.py:108-129), and no directories are created when ``dry_run`` is set (it is the default here)."""
import os
from datetime import datetime

from . import Networks as Nets
from . import data as DataHandeling

ROOT_DATA_DIR = '~/CellTrackingChallenge/Training/'
ROOT_TEST_DATA_DIR = '~/CellTrackingChallenge/Test/'
ROOT_SAVE_DIR = '~/LSTM-UNet-Outputs/'


def _with_defaults(**defaults):
    """Class decorator: the reference declares its configuration as class attributes (Params.py:28-95, 153-175); the
    same names and values are attached here from one table per class."""
    def attach(cls):
        for name, value in defaults.items():
            setattr(cls, name, value)
        return cls
    return attach


def _ctc_net_kernel_params():
    """Params.py:49-69: 3x3 conv pairs and one 5x5 ConvLSTM per encoder level, 3x3 conv pairs + the 1x1 logits conv in
    the decoder."""
    enc, dec = (128, 256, 256, 512), (256, 128, 64, 32)
    up = [[(3, f), (3, f)] for f in dec]
    up[-1].append((1, 3))
    return {'down_conv_kernels': [[(3, f), (3, f)] for f in enc], 'lstm_kernels': [[(5, f)] for f in enc],
            'up_conv_kernels': up}


class ParamsBase(object):
    aws = False

    def _override_params_(self, params_dict: dict):
        """Params.py:16-25: every key becomes an instance attribute; unknown keys are reported, not rejected."""
        known = set()
        for klass in type(self).__mro__:
            known.update(vars(klass))
        for key in params_dict:
            if key not in known:
                print('Warning!: Parameter:{} not in defualt parameters'.format(key))
            setattr(self, key, params_dict[key])


@_with_defaults(
    # general / data (Params.py:30-46)
    experiment_name='MyRun_SIM', gpu_id=0, data_provider_class=DataHandeling.SyntheticSequenceProvider,
    root_data_dir=ROOT_DATA_DIR, train_sequence_list=[('Fluo-N2DH-SIM+', '01'), ('Fluo-N2DH-SIM+', '02')],
    val_sequence_list=[('Fluo-N2DH-SIM+', '01'), ('Fluo-N2DH-SIM+', '02')], crop_size=(128, 128), batch_size=5,
    unroll_len=4, data_format='NCHW', train_q_capacity=200, val_q_capacity=200, num_val_threads=2, num_train_threads=8,
    # network (Params.py:48-69)
    net_model=Nets.ULSTMnet2D, net_kernel_params=_ctc_net_kernel_params(),
    # training (Params.py:71-76)
    class_weights=[0.15, 0.25, 0.6], learning_rate=1e-5, num_iterations=1000000, validation_interval=1000,
    print_to_console_interval=10,
    # save / restore, TensorBoard (Params.py:78-91)
    load_checkpoint=False, load_checkpoint_path='', continue_run=False, save_checkpoint_dir=ROOT_SAVE_DIR,
    save_checkpoint_iteration=5000, save_checkpoint_every_N_hours=24, save_checkpoint_max_to_keep=5,
    tb_sub_folder='LSTMUNet', write_to_tb_interval=500, save_log_dir=ROOT_SAVE_DIR,
    # debugging (the reference defaults dry_run to False; nothing is written here unless asked)
    dry_run=True, profile=False,
    # B200 backend (not in the reference)
    precision='bf16', seed=None)
class CTCParams(ParamsBase):
    def __init__(self, params_dict=None):
        self._override_params_(params_dict or {})
        if isinstance(self.gpu_id, list):
            os.environ['CUDA_VISIBLE_DEVICES'] = str(self.gpu_id)[1:-1]
        elif int(self.gpu_id) >= 0:
            os.environ.setdefault('CUDA_VISIBLE_DEVICES', str(self.gpu_id))
        self.train_data_base_folders = [(os.path.join(ROOT_DATA_DIR, ds[0]), ds[1]) for ds in self.train_sequence_list]
        self.val_data_base_folders = [(os.path.join(ROOT_DATA_DIR, ds[0]), ds[1]) for ds in self.val_sequence_list]
        self._train_provider = self._val_provider = None
        now_string = datetime.now().strftime('%Y-%m-%d_%H%M%S')
        if self.load_checkpoint and self.continue_run:        # Params.py:132-142: keep writing into the run being continued
            path = self.load_checkpoint_path
            if os.path.isdir(path):
                if path.endswith('tf-ckpt') or path.endswith('tf-ckpt/'):
                    self.experiment_log_dir = self.experiment_save_dir = os.path.dirname(path)
                else:
                    self.experiment_log_dir = self.experiment_save_dir = path
            else:
                self.experiment_log_dir = self.experiment_save_dir = os.path.dirname(os.path.dirname(path))
                self.load_checkpoint_path = os.path.join(path, 'tf-ckpt')
        else:
            self.experiment_log_dir = os.path.join(self.save_log_dir, self.tb_sub_folder, self.experiment_name, now_string)
            self.experiment_save_dir = os.path.join(self.save_checkpoint_dir, self.tb_sub_folder, self.experiment_name,
                                                    now_string)
        if not self.dry_run:
            os.makedirs(os.path.expanduser(self.experiment_save_dir), exist_ok=True)
        self.channel_axis = 1 if self.data_format == 'NCHW' else 3

    def _provider(self, folders, q_capacity, threads, seed):
        kw = dict(sequence_folder_list=folders, image_crop_size=self.crop_size, unroll_len=self.unroll_len, deal_with_end=0,
                  batch_size=self.batch_size, queue_capacity=q_capacity, data_format=self.data_format, randomize=True,
                  return_dist=False, num_threads=threads)
        # the reference seeds from OS entropy (no seed argument, Params.py:108-129).  `seed` (a params_dict entry, default
        # None) makes a run repeatable; in data-parallel training every rank gets its own stream (seed + rank), otherwise the
        # ranks would draw identical crops and average identical gradients.  Passed only to providers that accept it.
        import inspect
        base = getattr(self, 'seed', None)
        try:
            accepts = 'seed' in inspect.signature(self.data_provider_class).parameters
        except (TypeError, ValueError):
            accepts = False
        if accepts and base is not None:
            rank = 0
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized():
                    rank = dist.get_rank()
            except Exception:
                pass
            kw['seed'] = int(base) + 2 * rank + seed
        return self.data_provider_class(**kw)

    @property
    def train_data_provider(self):
        if self._train_provider is None:
            self._train_provider = self._provider(self.train_data_base_folders, self.train_q_capacity,
                                                  self.num_train_threads, 0)
        return self._train_provider

    @property
    def val_data_provider(self):
        if self._val_provider is None:
            self._val_provider = self._provider(self.val_data_base_folders, self.train_q_capacity,
                                                self.num_val_threads, 1)
        return self._val_provider


@_with_defaults(
    gpu_id=0, model_path='./Models/LSTMUNet2D/PhC-C2DL-PSC/', output_path='./tmp/output/PhC-C2DL-PSC/01',
    sequence_path=os.path.join(ROOT_TEST_DATA_DIR, 'PhC-C2DL-PSC/01/'), filename_format='t*.tif',
    data_reader=DataHandeling.CTCInferenceReader, data_format='NCHW',
    # instance labelling (Params.py:164-168; read by postprocess.PostProcessor)
    FOV=0, min_cell_size=10, max_cell_size=100, edge_dist=2, pre_sequence_frames=4,
    dry_run=True, save_intermediate=False, save_intermediate_path='./tmp/output/PhC-C2DL-PSC/01', precision='bf16')
class CTCInferenceParams(ParamsBase):
    def __init__(self, params_dict: dict = None):
        if params_dict is not None:
            self._override_params_(params_dict)
        self.channel_axis = 1 if self.data_format == 'NCHW' else 3
        if not self.dry_run:
            os.makedirs(self.output_path, exist_ok=True)
