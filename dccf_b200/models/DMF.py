"""`--n_layers` holder between RecModel and DCCF (src/models/DMF.py:11-24).  DMF's own cosine-MLP
scorer (DMF.py:37-82) is not on the DCCF path and is not provided."""
from .RecModel import RecModel


class DMF(RecModel):
    @staticmethod
    def parse_model_args(parser, model_name='DMF'):
        parser.add_argument('--n_layers', type=int, default=1, help='Number of mlp layers.')
        return RecModel.parse_model_args(parser, model_name)

    def __init__(self, label_min, label_max, feature_num, user_num, item_num, u_vector_size, i_vector_size,
                 n_layers, random_seed, model_path):
        self.n_layers = n_layers
        RecModel.__init__(self, label_min=label_min, label_max=label_max, feature_num=feature_num,
                          user_num=user_num, item_num=item_num, u_vector_size=u_vector_size,
                          i_vector_size=i_vector_size, random_seed=random_seed, model_path=model_path)
