/* TEST INFRASTRUCTURE ONLY: see ../postgres.h */
