"""Engine adapter with the reference's interface (engines/common_mil.py:1-69): dispatches one bag to the model entry
points and returns the 7-tuple (logits, label, aux_loss, patch_num, keep_num, pad_ratio, kn_std).  Works unchanged
with the reference's own CommonMIL as well -- the mhimk modules keep every forward() signature it calls."""


class CommonMIL:
    def __init__(self, args) -> None:
        self.training = True

    def init_func_train(self, args, **kwargs):
        self.training = True

    def init_func_val(self, args, **kwargs):
        self.training = False

    def after_get_data_func(self, args, **kwargs):
        pass

    def after_backward_func(self, args, **kwargs):
        pass

    def final_train_func(self, args, **kwargs):
        pass

    def forward_func(self, args, model, model_ema, bag, label, criterion, batch_size, i, epoch, n_iter, pos, **kwargs):
        pad_ratio, kn_std = 0.0, 0.0
        if args.model == "mhim":
            cls_tea, attn = model_ema.forward_teacher(bag) if model_ema is not None else (None, None)
            if args.aux_alpha == 0.0:
                cls_tea = None
            if args.baseline == "dsmil":
                tea = cls_tea[0] if cls_tea is not None else None        # (upstream indexes None here when aux_alpha == 0)
                logits, aux_loss, patch_num, keep_num = model(bag, attn, tea, i=n_iter)
                logits = 0.5 * logits[0].view(batch_size, -1) + 0.5 * logits[1].view(batch_size, -1)
            else:
                logits, aux_loss, patch_num, keep_num = model(bag, attn, cls_tea, i=n_iter)
        elif args.model == "mhim_pure":
            logits, aux_loss, patch_num, keep_num = model.pure(bag)
            if args.baseline == "dsmil":
                logits = 0.5 * logits[0].view(batch_size, -1) + 0.5 * logits[1].view(batch_size, -1)
        elif args.model in ("clam_sb", "clam_mb", "dsmil"):
            logits, aux_loss, _ = model(bag, label=label, loss=criterion, pos=pos)
            keep_num = patch_num = bag.size(1)
        else:
            logits = model(bag, pos=pos)
            aux_loss, patch_num, keep_num = 0.0, bag.size(1), bag.size(1)
        return logits, label, aux_loss, patch_num, keep_num, pad_ratio, kn_std

    def validate_func(self, args, model, bag, label, criterion, batch_size, i, pos, epoch=None, **kwargs):
        if args.model in ("mhim", "mhim_pure"):
            logits = model.forward_test(bag)
            if args.baseline == "dsmil":
                logits = logits[0]
        elif args.model == "dsmil":
            logits, _ = model(bag, pos=pos, epoch=epoch)
        else:
            logits = model(bag, pos=pos, epoch=epoch)
        if (args.model == "mhim" and isinstance(logits, (list, tuple))) or (args.model == "mhim_pure" and args.baseline == "dsmil"):
            logits = 0.5 * logits[0] + 0.5 * logits[1]
        return logits, label
