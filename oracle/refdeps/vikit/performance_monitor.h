// Stand-in for rpg_vikit performance_monitor.h — only the type name is needed.
#ifndef PLSVO_REFDEPS_VIKIT_PERFMON
#define PLSVO_REFDEPS_VIKIT_PERFMON
namespace vk {
class PerformanceMonitor {};
}
#endif
