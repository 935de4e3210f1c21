// Reward of the example custom environment `quadpod_walk` (contract:
// include/dial_custom_reward.h).  Runs on one lane per sample, once per env step.
//
// user[] (QuadpodEnv.user_params): 0 vx target, 1 ramp-up time, 2 torso height target,
// 3 energy weight, 4 foot-clearance weight, 5 max |ctrl|
DIAL_REWARD_FN float dial_custom_reward(const dial_reward_ctx* c) {
  const float* u = c->user;
  const int torso = 1;
  // forward velocity tracking, command ramped like the built-in walking envs
  const float t = (float)c->step * c->dt;
  const float vx_tar = fminf(u[0] * t / u[1], u[0]);
  float v[3], w[3];
  dial_xd_vel(c, torso, v);
  dial_xd_ang(c, torso, w);
  const float r_vel = -((v[0] - vx_tar) * (v[0] - vx_tar) + v[1] * v[1]);
  const float r_yawrate = -w[2] * w[2];
  // height and uprightness of the torso
  const float dz = c->xpos[3 * torso + 2] - u[2];
  const float* q = c->xquat + 4 * torso;
  const float upz = 1.f - 2.f * (q[1] * q[1] + q[2] * q[2]);   // z component of R e_z
  const float r_up = -(1.f - upz);
  // actuation effort and feet: stay close to the ground plane (contact.dist), under the hips
  float r_energy = 0.f;
  for (int a = 0; a < c->nu; ++a) { const float e = c->ctrl[a] / u[5]; r_energy -= e * e; }
  float r_feet = 0.f;
  for (int f = 0; f < c->ncon; ++f) r_feet -= c->contact_dist[f] * c->contact_dist[f];
  float head[3];
  dial_site_xpos(c, 0, head);
  const float r_head = -(head[2] - (u[2] + 0.02f)) * (head[2] - (u[2] + 0.02f));
  return r_vel + 0.1f * r_yawrate - 10.f * dz * dz + 0.5f * r_up + u[3] * r_energy + u[4] * r_feet + r_head;
}
