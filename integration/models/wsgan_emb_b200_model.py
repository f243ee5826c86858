"""Reference-side binding: copy (or put on `models.__path__`) as models/wsgan_emb_b200_model.py of phymhan/pc-gan and run
`train.py --model wsgan_emb_b200 ...`.  models/__init__.py:5-39 imports `models.<name>_model` and picks the class whose
lower-cased name is `<name>model` and that subclasses the reference's BaseModel, hence the second base."""
from models.base_model import BaseModel
from pcgan_b200.wsgan_emb_model import WSGANEmbModel as _B200


class WSGANEmbB200Model(_B200, BaseModel):
    def name(self):
        return "WSGANEmbB200Model"
