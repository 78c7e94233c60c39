"""Mirror of op3/torch/op3_modules/physics_network.py: registry `Physics_Args` (:12-29) and PhysicsNetwork (:33-187).
All eleven Mlps run inside op3_dynamics_fwd/bwd (csrc/dynamics.cu)."""
import torch
from torch import nn

from op3_b200.torch.networks import Mlp, identity

Physics_Args = dict(
    reg_ac32=("reg", lambda det_size, sto_size, action_size: dict(
        det_size=det_size, sto_size=sto_size, action_size=action_size, action_enc_size=32, hidden_activation=nn.ELU())),
    reg_ac32_no_share=("reg_no_share", lambda det_size, sto_size, action_size, k: dict(
        det_size=det_size, sto_size=sto_size, action_size=action_size, action_enc_size=32, hidden_activation=nn.ELU(), k=k)),
)


class PhysicsNetwork(nn.Module):
    def __init__(self, det_size, sto_size, action_size, action_enc_size, hidden_activation=nn.ELU(),
                 deterministic_state_activation=identity, lambda_output_activation=identity, k=None):
        super().__init__()
        if not isinstance(hidden_activation, nn.ELU) or action_enc_size != 32:
            raise NotImplementedError("only the 'reg_ac32' dynamics architecture (ELU, action_enc_size=32) is built")
        if deterministic_state_activation is not identity or lambda_output_activation is not identity:
            raise NotImplementedError("non-identity output activations are not built")
        self.K = k
        self.det_size, self.sto_size = det_size, sto_size
        self.full_rep_size = R = det_size + sto_size
        self.action_size = action_size
        self.action_enc_size = action_enc_size if action_size > 0 else 0
        self.effect_size = action_enc_size
        self.interaction_size = R
        act = hidden_activation
        mk = lambda hid, out, inp, out_act: Mlp((hid,), out, inp, hidden_activation=act, output_activation=out_act)
        self.inertia_encoder = mk(R, R, R, act)
        if action_size > 0:
            self.action_encoder = mk(R, self.action_enc_size, action_size, act)
            self.action_effect_network = mk(R, R, self.action_enc_size + R, act)
            self.action_attention_network = mk(R, 1, self.action_enc_size + R, nn.Sigmoid())
        self.pairwise_encoder_network = mk(2 * R, R, 2 * R, act)
        self.interaction_effect_network = mk(R, self.effect_size, R, act)
        self.interaction_attention_network = mk(R, 1, R, nn.Sigmoid())
        self.final_merge_network = mk(R, R, self.effect_size + R, act)
        if det_size != 0:
            self.det_output = mk(R, det_size, R, identity)
        self.lambdas1_output = mk(R, sto_size, R, identity)
        self.lambdas2_output = mk(R, sto_size, R, identity)

    def set_k(self, k):
        self.K = k

    def forward(self, sampled_state, actions):
        """sampled_state (B*K,R), actions (B*K,A) or None -> (deter (B*K,Rd) or None, lambdas1, lambdas2).
        Inference-only standalone entry point; training goes through OP3Model.run_schedule."""
        from op3_b200.torch.op3_modules._standalone import physics_forward
        return physics_forward(self, sampled_state, actions)[:3]

    def get_state_action_attention_values(self, sampled_state, actions):
        from op3_b200.torch.op3_modules._standalone import physics_forward
        K = self.K
        self.K = 1                      # the reference evaluates this per row, without pair interactions
        try:
            return physics_forward(self, sampled_state, actions)[3]
        finally:
            self.K = K

    def get_all_attention_values(self, sampled_state, actions, K):
        from op3_b200.torch.op3_modules._standalone import physics_forward
        old = self.K
        self.K = K
        try:
            _, _, _, act_att, pair_att, deltas = physics_forward(self, sampled_state, actions)
        finally:
            self.K = old
        return act_att, pair_att, deltas


class PhysicsNetwork_v2_No_Sharing(nn.Module):
    """physics_network.py:192-249: K full PhysicsNetworks (`models.<k>...` state_dict keys).  Model k is run on ALL rows
    (every slot interacts with every other through model k's weights) and only its rows (b, k) are kept; the K results are
    concatenated slot-major (output row k*bs + b), the reference's row order."""
    no_sharing = True

    def __init__(self, det_size, sto_size, action_size, action_enc_size, hidden_activation=nn.ELU(),
                 deterministic_state_activation=identity, lambda_output_activation=identity, k=None):
        super().__init__()
        if k is None:
            raise ValueError("A value of k is needed to initialize this model!")
        self.K = k
        self.det_size, self.sto_size, self.action_size = det_size, sto_size, action_size
        self.models = nn.ModuleList([
            PhysicsNetwork(det_size, sto_size, action_size, action_enc_size, hidden_activation, deterministic_state_activation,
                           lambda_output_activation, k=k) for _ in range(k)])

    def forward(self, sampled_state, actions):
        """sampled_state (B*K,R), actions (B*K,A) or None -> (deter (B*K,Rd) or None, lambdas1, lambdas2), rows slot-major."""
        det, l1, l2 = [], [], []
        for i in range(self.K):
            vals = self.models[i](sampled_state, actions)
            det.append(self._get_ith_input(vals[0], i))
            l1.append(self._get_ith_input(vals[1], i))
            l2.append(self._get_ith_input(vals[2], i))
        return (None if det[0] is None else torch.cat(det)), torch.cat(l1), torch.cat(l2)

    def _get_ith_input(self, x, i):
        return None if x is None else x.view([-1, self.K] + list(x.shape[1:]))[:, i]

    def set_k(self, k):
        assert k == self.K
