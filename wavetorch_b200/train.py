"""Training loop of the reference with the data path kept on the GPU.

`train(...)` has the signature, the epoch structure and the history columns of wavetorch/train.py:13-133:
epoch 0 only characterises the starting structure, epochs 1..N take one optimizer step per batch followed by
`constrain_to_design_region()`, then a no-grad pass over the training set (accuracy, confusion matrix) and one over the
test set (loss, accuracy, confusion matrix).  Differences, all on the host side of the path:
  * batches are moved to the model's device once per use with non_blocking copies (the reference leaves placement to the
    caller and then calls `.numpy()` on what may be CUDA tensors, train.py:93,106);
  * `criterion` instances of torch.nn.CrossEntropyLoss with default options run as the fused head
    (wavetorch_b200.loss.power_cross_entropy: one launch, analytic backward seed); any other criterion goes through
    `criterion(normalize_power(out.sum(dim=1)), labels)` exactly as train.py:61-62;
  * evaluation passes run under torch.no_grad(), which on this path also means no adjoint tape is written;
  * per-batch losses are collected as device scalars and read back once per epoch, not with `.item()` per batch;
  * pandas >= 2 has no DataFrame.append (train.py:116): rows are concatenated.
"""
import copy
import time

import numpy as np
import torch

from .io import save_model
from .loss import power_cross_entropy
from .utils import accuracy_onehot, normalize_power


def confusion_matrix(y_true, y_pred):
    """sklearn.metrics.confusion_matrix for integer label tensors: rows = truth, columns = prediction, over the sorted
    union of the labels that occur (train.py:93,106 call it without `labels=`)."""
    y_true = torch.as_tensor(y_true).reshape(-1).cpu().to(torch.int64)
    y_pred = torch.as_tensor(y_pred).reshape(-1).cpu().to(torch.int64)
    labels = torch.unique(torch.cat([y_true, y_pred]))
    n = labels.numel()
    ti = torch.searchsorted(labels, y_true)
    pi = torch.searchsorted(labels, y_pred)
    cm = torch.bincount(ti * n + pi, minlength=n * n).reshape(n, n)
    return cm.numpy()


def _is_plain_cross_entropy(criterion):
    if not isinstance(criterion, torch.nn.CrossEntropyLoss):
        return False
    return (criterion.weight is None and criterion.reduction == "mean" and criterion.ignore_index == -100
            and getattr(criterion, "label_smoothing", 0.0) == 0.0)


def _head(model, criterion, xb, labels, fused):
    out = model(xb)
    if fused:
        return power_cross_entropy(out, labels)
    yb_pred = normalize_power(out.sum(dim=1))
    return criterion(yb_pred, labels), yb_pred


def train(model, optimizer, criterion, train_dl, test_dl, N_epochs: int, batch_size: int, history=None,
          history_model_state=[], fold=None, name=None, savedir=None, cfg=None, accuracy=None):
    """Trains the model; returns (history DataFrame, history_model_state) like wavetorch.train (train.py:13-133)."""
    import pandas as pd
    inner = getattr(model, "model", model)
    dev = next(inner.parameters()).device
    fused = _is_plain_cross_entropy(criterion)
    if history is None:
        history = pd.DataFrame(columns=['time', 'epoch', 'fold', 'loss_train', 'loss_test', 'acc_train', 'acc_test',
                                        'cm_train', 'cm_test'])

    def to_dev(t):
        return t.to(dev, non_blocking=True)

    t_start = time.time()
    for epoch in range(0, N_epochs + 1):
        t_epoch = time.time()
        loss_iter = []
        for num, (xb, yb) in enumerate(train_dl):
            xb, labels = to_dev(xb), to_dev(yb).argmax(dim=1)

            def closure():
                optimizer.zero_grad()
                loss, _ = _head(model, criterion, xb, labels, fused)
                loss.backward()
                return loss

            if epoch == 0:      # don't take a step, just characterise the starting structure
                with torch.no_grad():
                    loss, _ = _head(model, criterion, xb, labels, fused)
            else:
                loss = optimizer.step(closure)
                inner.cell.geom.constrain_to_design_region()
            loss_iter.append(loss.detach())

        with torch.no_grad():
            acc_train_tmp, list_pred, list_truth = [], [], []
            for num, (xb, yb) in enumerate(train_dl):
                xb, yb = to_dev(xb), to_dev(yb)
                yb_pred = normalize_power(model(xb).sum(dim=1))
                list_pred.append(yb_pred)
                list_truth.append(yb)
                if accuracy is not None:
                    acc_train_tmp.append(accuracy(yb_pred, yb.argmax(dim=1)))
            cm_train = confusion_matrix(torch.cat(list_truth).argmax(dim=1), torch.cat(list_pred).argmax(dim=1))

            acc_test_tmp, loss_test_tmp, list_pred, list_truth = [], [], [], []
            cm_test = None
            if test_dl is not None:
                for num, (xb, yb) in enumerate(test_dl):
                    xb, yb = to_dev(xb), to_dev(yb)
                    labels = yb.argmax(dim=1)
                    loss, yb_pred = _head(model, criterion, xb, labels, fused)
                    list_pred.append(yb_pred)
                    list_truth.append(yb)
                    loss_test_tmp.append(loss.detach())
                    if accuracy is not None:
                        acc_test_tmp.append(accuracy_onehot(yb_pred, labels))
                cm_test = confusion_matrix(torch.cat(list_truth).argmax(dim=1), torch.cat(list_pred).argmax(dim=1))

        mean = lambda xs: float(torch.stack([torch.as_tensor(v, dtype=torch.float64).cpu() for v in xs]).mean()) if len(xs) else float("nan")
        loss_train, loss_test = mean(loss_iter), mean(loss_test_tmp)
        acc_train = float(np.mean(acc_train_tmp)) if acc_train_tmp else float("nan")
        acc_test = float(np.mean(acc_test_tmp)) if acc_test_tmp else float("nan")
        print('Epoch %2d/%2d --- Elapsed Time:  %4.2f min | Training Loss:  %.4e | Testing Loss:  %.4e | Training Accuracy:  %.4f | Testing Accuracy:  %.4f'
              % (epoch, N_epochs, (time.time() - t_epoch) / 60, loss_train, loss_test, acc_train, acc_test))
        row = {'time': pd.to_datetime('now'), 'epoch': epoch, 'fold': fold, 'loss_train': loss_train,
               'loss_test': loss_test, 'acc_train': acc_train, 'acc_test': acc_test, 'cm_train': cm_train,
               'cm_test': cm_test}
        history = pd.concat([history, pd.DataFrame([row])], ignore_index=True)
        history_model_state.append(copy.deepcopy(inner.cell.geom.state_reconstruction_args()))
        if name is not None:
            save_model(model, name, savedir, history, history_model_state, cfg, verbose=False)

    print('Total Time: %.2f min\n' % ((time.time() - t_start) / 60))
    return history, history_model_state
