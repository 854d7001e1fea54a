"""Embedding-model base of the reference's protocol (src/models/RecModel.py:9-36): 'X' carries the user and item id
only (`append_id`, nothing else), `--u_vector_size` / `--i_vector_size` (equal, default 64), one embedding table per
side.  Its own `predict` is the plain matrix-factorisation dot product; DCCF overrides everything below the flags."""
import torch

from .BaseModel import BaseModel


class RecModel(BaseModel):
    # how DataProcessor.format_data_dict lays out 'X' for this family: [uid, iid], no side-feature columns
    append_id, include_id = True, False
    include_user_features = include_item_features = False

    @staticmethod
    def parse_model_args(parser, model_name='RecModel'):
        for side, what in (('u', 'user'), ('i', 'item')):
            parser.add_argument('--%s_vector_size' % side, type=int, default=64, help='Size of %s vectors.' % what)
        return BaseModel.parse_model_args(parser, model_name)

    def __init__(self, label_min, label_max, feature_num, user_num, item_num, u_vector_size, i_vector_size,
                 random_seed, model_path):
        if u_vector_size != i_vector_size:
            raise AssertionError('u_vector_size (%s) and i_vector_size (%s) must be equal' % (u_vector_size, i_vector_size))
        # plain attributes first: BaseModel.__init__ seeds the generators and calls _init_weights, which needs them
        self.user_num, self.item_num = user_num, item_num
        self.u_vector_size = self.i_vector_size = self.ui_vector_size = u_vector_size
        super().__init__(label_min, label_max, feature_num, random_seed=random_seed, model_path=model_path)

    def _init_weights(self):
        # user table before item table: the order in which the torch generator is consumed is part of the contract
        for name, rows in (('uid_embeddings', self.user_num), ('iid_embeddings', self.item_num)):
            setattr(self, name, torch.nn.Embedding(rows, self.ui_vector_size))

    def predict(self, feed_dict):
        ids = feed_dict['X']
        score = (self.uid_embeddings(ids[:, 0]) * self.iid_embeddings(ids[:, 1])).sum(dim=1)
        return {'prediction': score.view([-1]), 'check': []}
