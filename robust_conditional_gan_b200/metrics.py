"""Evaluation metrics of the reference: generated-label accuracy under a frozen classifier
(cifar10/gan_resnet.py:424-455, mnist/utils.py:273-306), inception score (cifar10/common/inception/inception_score_.py), label
recovery zero-one loss (mnist/model.py:626-638).

The reference evaluates with frozen TensorFlow GraphDefs (`resnet-110/graph_optimized.pb`, `mnist_dcnn/graph_optimized.pb`,
the TF-GAN inception graph) that are neither in the repository nor obtainable offline, so the classifier is a parameter here:
any callable `images -> class probabilities / predictions`.  Everything downstream of the classifier -- label mapping through
the learned confusion matrix, per-class batching, split statistics -- follows the reference line by line."""
import numpy as np


def preds2score(preds, splits=10):
    """inception_score_.py:56-63: exp(E_x KL(p(y|x) || p(y))) per split -> (mean, std)."""
    preds = np.asarray(preds, dtype=np.float64)
    scores = []
    for i in range(splits):
        part = preds[(i * preds.shape[0] // splits):((i + 1) * preds.shape[0] // splits), :]
        kl = part * (np.log(part) - np.log(np.expand_dims(np.mean(part, 0), 0)))
        kl = np.mean(np.sum(kl, 1))
        scores.append(np.exp(kl))
    return float(np.mean(scores)), float(np.std(scores))


def softmax(logits):
    """inception_score_.py:52: np.exp(preds) / np.sum(np.exp(preds), 1, keepdims=True) (max-shifted for stability)"""
    z = np.asarray(logits, dtype=np.float64)
    z = np.exp(z - z.max(1, keepdims=True))
    return z / z.sum(1, keepdims=True)


def get_inception_score(images, logits_fn, splits=10, batch_size=128):
    """inception_score_.py:44-75 with the network as a parameter: images [N,3,H,W] in [-1,1]; logits_fn(batch) -> [b, >=1000]
    logits (the reference keeps the first 1000 columns); whole batches only, like the reference."""
    images = np.asarray(images)
    assert images.ndim == 4 and images.shape[1] == 3 and np.max(images[0]) <= 1 and np.min(images[0]) >= -1
    preds = []
    for i in range(len(images) // batch_size):
        preds.append(np.asarray(logits_fn(images[i * batch_size:(i + 1) * batch_size]))[:, :1000])
    return preds2score(softmax(np.concatenate(preds, 0)), splits)


def map_labels_through_confusion(labels, confusion_matrix, vocab_size=10):
    """gan_resnet.py:430-440: with a learned confusion matrix the generator's class k may correspond to the true class
    argmax_j C[k, j]; labels are sent through that hard permutation before comparing."""
    C = np.asarray(confusion_matrix)
    hard = np.zeros_like(C, dtype=int)
    hard[np.arange(C.shape[0]), np.argmax(C, axis=-1)] = 1
    onehot = np.zeros([len(labels), vocab_size], dtype=float)
    onehot[np.arange(len(labels)), np.asarray(labels, dtype=int)] = 1
    return np.argmax(onehot.dot(hard), axis=-1)


def generated_label_accuracy(samples, labels, classify, confusion_matrix=None):
    """gan_resnet.py:424-455: fraction of generated samples whose frozen-classifier prediction equals the label they were
    generated for.  classify(samples) -> [N, classes] scores (the reference: softmax of a pre-trained ResNet-110)."""
    labels = np.asarray(labels)
    if confusion_matrix is not None:
        labels = map_labels_through_confusion(labels, confusion_matrix)
    scores = np.asarray(classify(samples))
    return float((labels == np.argmax(scores, axis=-1)).astype(float).mean())


def generated_label_accuracy_mnist(samples, predict, num_test=100):
    """mnist/utils.py:273-306: samples [100 sampler runs, 100 per run, 28, 28, 1] where within each run the i-th group of 10
    shares label i (mnist/model.py:458-471; the reference's reshape only works for 100 runs of 100), regrouped per class,
    classified in batches of num_test, per-batch accuracies averaged.  predict(x) -> predicted class ids."""
    samples = np.asarray(samples)
    test_images = (samples.transpose((1, 0, 2, 3, 4)).reshape((10, 10) + samples.shape[1:]).reshape((10, -1) + samples.shape[2:]))
    acc_sum, num_sum = 0.0, 0
    for y_actual, class_samples in enumerate(test_images):
        for ii in range(num_test, class_samples.shape[0] + 1, num_test):
            y = np.asarray(predict(class_samples[ii - num_test:ii]))
            acc_sum += float((y == y_actual).astype(float).mean())
            num_sum += 1
    return acc_sum / num_sum


def zero_one_loss(y_actual, y_recover):
    """mnist/model.py:626-632: mean(1 - sum(y_actual * onehot(argmax y_recover)))."""
    y_actual, y_recover = np.asarray(y_actual, dtype=np.float64), np.asarray(y_recover)
    hard = np.eye(y_actual.shape[1])[np.argmax(y_recover, axis=-1)]
    return float(np.mean(1.0 - np.sum(y_actual * hard, axis=-1)))
