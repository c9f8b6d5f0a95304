/* pup.h -- stand-in for Charm++'s header of that name: MultipoleMoments.h only declares a friend
 * operator|(PUP::er &, MultipoleMoments &) whose definition sits behind __CHARMC__.  TEST INFRASTRUCTURE ONLY. */
#ifndef CB200_ORACLE_SHIM_PUP_H
#define CB200_ORACLE_SHIM_PUP_H
namespace PUP {
class er;
}
#endif
