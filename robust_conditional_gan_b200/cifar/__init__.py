"""Drop-ins for cifar10/gan_resnet.py and cifar10/common/ops/* on the B200 static-program runtime."""
