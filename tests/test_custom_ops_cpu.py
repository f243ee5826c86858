"""The torch.library operators of the hot path (pcgan_b200/custom_ops.py) are registered with schemas, fake (meta) kernels
and autograd formulas: checked here without a GPU through FakeTensor shape inference."""
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from pcgan_b200 import networks as NW

OPS = ["resnet_generator", "resnet_generator_backward", "unet_generator", "alexnet_feature", "alexnet_feature_backward", "nlayer_discriminator", "nlayer_discriminator_backward", "siamese_feature",
       "siamese_feature_backward", "reduce_loss", "reduce_loss_backward", "upsample_bilinear_ac", "upsample_bilinear_ac_backward"]


def test_operators_are_registered_with_schemas():
    for name in OPS:
        op = getattr(torch.ops.pcgan, name)
        schema = op.default._schema
        assert schema.name == "pcgan::" + name
    s = str(torch.ops.pcgan.resnet_generator.default._schema)
    assert "Tensor x, Tensor z, Tensor[] params" in s and s.endswith("-> Tensor")


def test_fake_tensor_shape_inference():
    g = NW.ResnetGenerator(3, 3, 1, 64, NW.get_norm_layer("instance"), n_blocks=2)
    d = NW.NLayerDiscriminator(3, 1, 64, 3, NW.get_norm_layer("batch"), True)
    e = NW.SiameseFeature(NW.ResNetFeature(), pooling="avg", cnn_dim=[32, 1], cnn_pad=1, cnn_relu_slope=0.7, noisy=True,
                          drop_layer=NW.get_dropout_layer(0.0))
    assert d.out_size(128) == 14 and d.in_size(14) == 128 and d.out_size(256) == 30
    with FakeTensorMode(allow_non_fake_inputs=True):
        x = torch.empty(2, 3, 128, 128, device="cuda")
        z = torch.empty(2, 1, 1, 1, device="cuda")
        out = torch.ops.pcgan.resnet_generator(x, z, [], g._key)
        assert tuple(out.shape) == (2, 3, 128, 128) and out.dtype == torch.float32 and out.device.type == "cuda"
        dx, dz = torch.ops.pcgan.resnet_generator_backward(out, out, g._key, True, True, False)
        assert tuple(dx.shape) == (2, 3, 128, 128) and tuple(dz.shape) == (2,)
        p = torch.ops.pcgan.nlayer_discriminator(x, z, [], d._key)
        assert tuple(p.shape) == (2, 1, 14, 14)
        dx, dz = torch.ops.pcgan.nlayer_discriminator_backward(p, p, d._key, True, False, True)
        assert tuple(dx.shape) == (2, 3, 128, 128) and dz.numel() == 0
        y, lv = torch.ops.pcgan.siamese_feature(torch.empty(2, 3, 224, 224, device="cuda"), [], e._key)
        assert tuple(y.shape) == (2, 1, 1, 1) and tuple(lv.shape) == (2, 1, 1, 1)
        loss = torch.ops.pcgan.reduce_loss(0, p, torch.empty(2, device="cuda"), 196)
        assert loss.dim() == 0
        gp = torch.ops.pcgan.reduce_loss_backward(loss, 0, p, torch.empty(2, device="cuda"), 196)
        assert gp.shape == p.shape
        up = torch.ops.pcgan.upsample_bilinear_ac(x, 224)
        assert tuple(up.shape) == (2, 3, 224, 224)
        assert tuple(torch.ops.pcgan.upsample_bilinear_ac_backward(up, 128, 128).shape) == (2, 3, 128, 128)
