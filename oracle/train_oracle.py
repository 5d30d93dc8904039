"""TEST INFRASTRUCTURE ONLY - CPU oracle for the training / evaluation side of the RADNet hot path
(SURVEY.md 8(f) f4): the four losses of `faster_rcnn/losses.py:16-95` and the mAP evaluation of
`test.py:48-173`.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

Parity pinning
--------------
* `get_objects`, `calc_class_ap` (test.py:48-173): PINNED - the reference functions run unmodified in the build
  container (extracted from test.py's source with `ast`, its script-level imports are absent here);
  `oracle/make_golden_f4.py` stores their outputs as tests/golden/f4_eval.npz.
* the losses: PARITY UNPINNED against Keras/TensorFlow (not installable here).  `losses.py` is a composition of
  Keras-2.2 / TF-1 backend calls whose definitions are public: K.abs, K.less_equal, K.cast, K.sum, K.mean,
  K.binary_crossentropy(target, output) = sigmoid_cross_entropy_with_logits(labels=target,
  logits=log(clip(output, 1e-7, 1-1e-7) / (1 - clip(...)))), and keras.objectives.categorical_crossentropy =
  -sum(target * log(clip(output / sum(output), 1e-7, 1-1e-7))).  The restatement evaluates every element in float32
  (K.floatx) and accumulates the sums in float64; TF's own reduction order is unspecified, so the bar is 1e-5
  relative (north_star tolerance), stated in the tests.  NOTE the reference calls
  K.binary_crossentropy(y_pred, y_true[..., A:]) - prediction in the `target` slot, label in the `output` slot
  (losses.py:65; same in upstream keras-frcnn) - and that is what is restated.
"""
import numpy as np

EPSILON = np.float32(1e-4)          # losses.py:14
K_EPS = np.float32(1e-7)            # keras.backend.epsilon()
F32 = np.float32


def _smooth_l1_sum(y_mask, x):
    """sum(mask * (x_bool*(0.5*x*x) + (1-x_bool)*(|x|-0.5))) / sum(eps + mask)  (losses.py:31-42, 81-84)."""
    x = x.astype(F32)
    x_abs = np.abs(x)
    x_bool = (x_abs <= F32(1.0)).astype(F32)
    term = y_mask * (x_bool * (F32(0.5) * x * x) + (F32(1.0) - x_bool) * (x_abs - F32(0.5)))
    num = term.astype(np.float64).sum()
    den = (EPSILON + y_mask).astype(np.float64).sum()
    return F32(F32(num) / F32(den))


def rpn_loss_regr(num_anchors):
    def f(y_true, y_pred):                                     # losses.py:31-42
        y_true = np.asarray(y_true).astype(F32)
        y_pred = np.asarray(y_pred).astype(F32)
        return _smooth_l1_sum(y_true[..., :4 * num_anchors], y_true[..., 4 * num_anchors:] - y_pred)
    return f


def _bce_swapped(target, output):
    """K.binary_crossentropy(target, output) of Keras 2.2 (TF backend), float32."""
    out = np.clip(output, K_EPS, F32(1.0) - K_EPS).astype(F32)
    logits = np.log(out / (F32(1.0) - out)).astype(F32)
    # tf.nn.sigmoid_cross_entropy_with_logits: max(x, 0) - x*z + log(1 + exp(-|x|))
    return (np.maximum(logits, F32(0)) - logits * target + np.log1p(np.exp(-np.abs(logits)))).astype(F32)


def rpn_loss_cls(num_anchors):
    def f(y_true, y_pred):                                     # losses.py:63-65
        y_true = np.asarray(y_true).astype(F32)
        y_pred = np.asarray(y_pred).astype(F32)
        valid = y_true[..., :num_anchors]
        term = valid * _bce_swapped(y_pred, y_true[..., num_anchors:])
        return F32(F32(term.astype(np.float64).sum()) / F32((EPSILON + valid).astype(np.float64).sum()))
    return f


def class_loss_regr(num_classes):
    def f(y_true, y_pred):                                     # losses.py:79-86
        y_true = np.asarray(y_true).astype(F32)
        y_pred = np.asarray(y_pred).astype(F32)
        return _smooth_l1_sum(y_true[..., :4 * num_classes], y_true[..., 4 * num_classes:] - y_pred)
    return f


def class_loss_cls(y_true, y_pred):                            # losses.py:93-95
    t = np.asarray(y_true)[0].astype(F32)
    o = np.asarray(y_pred)[0].astype(F32)
    o = o / o.sum(axis=-1, keepdims=True, dtype=F32)
    o = np.clip(o, K_EPS, F32(1.0) - K_EPS)
    rows = -(t * np.log(o)).astype(F32).astype(np.float64).sum(axis=-1)
    return F32(rows.mean())


# ------------------------------------------------------------------ evaluation (test.py:48-173)
def iou(a, b):
    """faster_rcnn/utils.py:77-109."""
    if a[0] >= a[2] or a[1] >= a[3] or b[0] >= b[2] or b[1] >= b[3]:
        return 0.0
    x = max(a[0], b[0]); y = max(a[1], b[1])
    w = min(a[2], b[2]) - x; h = min(a[3], b[3]) - y
    if w < 0 or h < 0:
        return 0.0
    inter = w * h
    union = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter
    return float(inter) / float(union + 1e-6)


def get_objects(pred, gt, threshold):
    """test.py:48-113: greedy matching of one image's detections (descending score; ties: higher index first, the
    reversed stable ascending argsort) against its figures.  Returns (T, P): per class, match flags and scores in
    visiting order, then one (1, 0) pair per unmatched figure."""
    T, P = {}, {}
    matched = [False] * len(gt)
    probs = np.array([s['prob'] for s in pred])
    order = np.argsort(probs, kind="stable")[::-1] if len(pred) else []
    for i in order:
        pb = pred[i]
        c = pb['class']
        if c not in P:
            P[c], T[c] = [], []
        P[c].append(pb['prob'])
        found = False
        for k, g in enumerate(gt):
            if g['class'] != c or matched[k]:
                continue
            if iou((pb['x1'], pb['y1'], pb['x2'], pb['y2']), (g['x1'], g['y1'], g['x2'], g['y2'])) >= threshold:
                found = True
                matched[k] = True
                break
        T[c].append(int(found))
    for k, g in enumerate(gt):
        if not matched[k]:
            if g['class'] not in P:
                P[g['class']], T[g['class']] = [], []
            T[g['class']].append(1)
            P[g['class']].append(0)
    return T, P


def calc_class_ap(y_true, y_pred):
    """test.py:117-173 (VOC-style AP of one class)."""
    y_true = np.array(y_true)
    y_pred = np.array(y_pred)
    n_gt = np.sum(y_true)
    order = np.flip(np.argsort(y_pred, kind="stable"))
    tp = fp = 0
    precision, recall = [], []
    for i in order:
        if y_true[i] > 0 and y_pred[i] > 0.0:
            tp += 1
        elif y_true[i] == 0 and y_pred[i] > 0.0:
            fp += 1
        precision.append(0.0 if tp + fp == 0 else tp / (tp + fp))
        recall.append(tp / n_gt if n_gt != 0 else 0.0)
    precision = np.array(precision)
    recall = np.array(recall)
    ip, ir = [], []
    mx = 0.0
    for i in reversed(range(len(recall))):
        if precision[i] > mx:
            mx = precision[i]
        ir.append(recall[i])
        ip.append(mx)
    ip.reverse()
    ir.reverse()
    ap = 0
    for i in range(len(ip) - 1):
        ap += ip[i + 1] * (ir[i + 1] - ir[i])
    return ap, precision, recall, ip, ir
