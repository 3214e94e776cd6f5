"""models.capsules_jhmdb_semi_sup_pa -- imported by the reference's main_jhmdb.py:338 / evaluate_jhmdb.py:25 but
ABSENT from the reference repository.  Reconstructed from the UCF model: 21 classes
(main_jhmdb.py:383 SpreadLoss(num_class=21)) => ConvCaps(32, 21, ...) and upsample1 with 336 input channels;
constructor signature CapsNet(pretrained_load=True) as called at main_jhmdb.py:371."""
from models.capsules_ucf101 import CapsNet as _UCFCapsNet
from models.capsules_ucf101 import ConvCaps, PrimaryCaps  # noqa: F401  (re-exported like the UCF module)


class CapsNet(_UCFCapsNet):
    NUM_CLASSES = 21

    def __init__(self, pt_path='../weights/rgb_charades.pt', P=4, pretrained_load=True):
        super(CapsNet, self).__init__(pt_path=pt_path, P=P, pretrained_load=pretrained_load)
