"""Mirror of the reference's ``losses.WeightedCELoss`` (losses.py:8-27) for the CUDA model.

The reference computes the loss from the logits tensor the model returned; here the logits of the last model call are
still resident in the library's workspace (fp32, exactly the values returned), so the loss kernel reads them there.
``seg_measure`` (losses.py:29-88, a CPU scipy metric) is out of scope."""


class WeightedCELoss(object):
    def __init__(self, channel_axis, class_weights):
        self.channel_axis = channel_axis
        self.class_weights = class_weights

    def __call__(self, gt_sequence, output_sequence):
        model = getattr(output_sequence, '_lu_model', None)
        if model is None:
            raise ValueError('WeightedCELoss expects the logits returned by the last ULSTMnet2D call')
        return model.loss(gt_sequence, self.class_weights)
