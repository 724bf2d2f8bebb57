function [plant_final, steps, n_hist, plant_hist, exit_hist] = main_closed_loop_b200(model, N_steps, dt, n_sim, target_vel, plant0, x_opt0, u_opt0)
%MAIN_CLOSED_LOOP_B200 The simulation loop of main.m:90-190 for B vehicles on the device: projection onto the track
%(cartesian_to_curvilinear.m, closest_point.m), x0 and the speed-ramp reference (main.m:92-114), the fused LTV-MPC
%step linearised at the previous prediction (main.m:118-127), actuator PIDs and the Cartesian dynamic plant
%(pid_controller.m, integrate_cart_dyn.m).  plant0 [7 x B], x_opt0 [N_x*N x B], u_opt0 [N_u*N x B] (main.m:44-55).
    h = fsae_mpc_b200_handle();
    [plant_final, steps, n_hist, plant_hist, exit_hist] = fsae_mpc_b200_mex('closed_loop', h, model, N_steps, dt, n_sim, ...
        target_vel, plant0, x_opt0, u_opt0);
end
