// Reward of the example custom environment `pincher_spin` (contract: include/dial_custom_reward.h):
// keep the ball where it is between the finger tips while spinning it about the vertical axis (the
// manipulation.py recipe of the reference: angular-velocity tracking + position + joint deviation).
//
// user[] (PincherEnv.user_params): 0 ball height, 1 target spin rate (rad/s about z), 2 joint-deviation
// weight, 3..6 nominal finger joint angles
DIAL_REWARD_FN float dial_custom_reward(const dial_reward_ctx* c) {
  const float* u = c->user;
  const int ball = 1;
  float w[3];
  dial_xd_ang(c, ball, w);
  const float r_spin = -((w[2] - u[1]) * (w[2] - u[1]) + w[0] * w[0] + w[1] * w[1]);
  const float* p = c->xpos + 3 * ball;
  const float r_pos = -(p[0] * p[0] + p[1] * p[1] + (p[2] - u[0]) * (p[2] - u[0]));
  float r_joint = 0.f;
  for (int j = 0; j < 4; ++j) { const float d = c->qpos[7 + j] - u[3 + j]; r_joint -= d * d; }
  // contacts 0 .. ncon-1 as modelc orders them; the closest fingertip-ball distance pulls the tips in
  float dmin = 1.f;
  for (int k = 0; k < c->ncon; ++k) dmin = fminf(dmin, c->contact_dist[k]);
  return 0.05f * r_spin + 50.f * r_pos + u[2] * r_joint - 2.f * fmaxf(dmin, 0.f);
}
