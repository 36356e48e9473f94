from .resnet import (ResNet, BasicBlock, Bottleneck, TinyConvNet, MODEL_ZOO,
                     resnet18, resnet34, resnet50, resnet101, resnet152,
                     init_imagenet_in_1hr)
