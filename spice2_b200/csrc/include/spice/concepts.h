// Compile-time plugin contract for user models — same names and meaning as the reference's
// spice/include/spice/concepts.h:11-104,163-228, so user structs written against the reference
// are checked identically here.  Deltas for the B200 backend:
//   * member functions that run inside kernels carry SPICE_HD;
//   * a neuron whose update() draws from the rng declares how many engine draws one call
//     consumes: `static constexpr int rng_draws = 1;` (needed for bit-exact random access into
//     the reference's one-stream-per-step RNG, snn.cpp:12-15).
#pragma once

#include <concepts>
#include <random>
#include <span>
#include <type_traits>
#include <vector>

#include "spice/util/platform.h"

namespace spice {
namespace util {
struct any_t {
	template <class T>
	constexpr operator T&() const;
};
template <bool... B>
inline constexpr int count_true = (0 + ... + int(B));
}

// ---- neurons (reference: concepts.h:11-57) ------------------------------------------------------
template <class T>
concept StatelessNeuron = std::default_initializable<T>;

template <class T>
concept StatefulNeuron = std::default_initializable<T> && requires { typename T::neuron; } &&
                         std::default_initializable<typename T::neuron>;

template <class T>
concept PerNeuronInit = StatefulNeuron<T> && requires(T const t, typename T::neuron& n, Int id, std::mt19937& rng) {
	t.init(n, id, rng);
};

template <class T>
concept PerPopulationInit = StatefulNeuron<T> && requires(T t, std::span<typename T::neuron> n, std::mt19937& rng) {
	t.init(n, rng);
};

namespace detail {
template <class T>
concept stateful_update = requires(T const t, typename T::neuron& n, float dt, std::mt19937& rng) {
	{ t.update(n, dt, rng) } -> std::same_as<bool>;
};
template <class T>
concept stateless_update = requires(T const t, float dt, std::mt19937& rng) {
	{ t.update(dt, rng) } -> std::same_as<bool>;
};
}

template <class T>
concept PerNeuronUpdate = (StatefulNeuron<T> && detail::stateful_update<T>) ||
                          (!StatefulNeuron<T> && detail::stateless_update<T>);

template <class T>
concept PerPopulationUpdate = std::default_initializable<T> &&
                              (std::copy_constructible<T> || std::move_constructible<T>) &&
                              requires(T t, float dt, std::mt19937& rng, std::vector<Int32>& out_spikes) {
	                              t.update(dt, rng, out_spikes);
                              };

template <class T>
concept Neuron = (PerPopulationUpdate<T> &&
                  util::count_true<StatefulNeuron<T>, PerNeuronUpdate<T>, PerNeuronInit<T>, PerPopulationInit<T>> == 0) ||
                 (!PerPopulationUpdate<T> && (StatelessNeuron<T> || StatefulNeuron<T>) && PerNeuronUpdate<T> &&
                  util::count_true<PerNeuronInit<T>, PerPopulationInit<T>> <= 1);

// ---- synapses (reference: concepts.h:59-104) ------------------------------------------------------
template <class T>
concept StatelessSynapse = std::default_initializable<T>;

template <class T>
concept StatefulSynapse = std::default_initializable<T> && requires { typename T::synapse; } &&
                          std::default_initializable<typename T::synapse>;

namespace detail {
template <class T, class N>
concept deliver_to_stateful = requires(T const t, typename T::synapse const& s, typename N::neuron& n) { t.deliver(s, n); };
template <class T, class N>
concept deliver_to_stateless = requires(T const t, typename N::neuron& n) { t.deliver(n); };
template <class T, class S, class D>
concept deliver_from_to_stateful = requires(T const t, typename T::synapse const& syn, typename S::neuron const& s,
                                            typename D::neuron& d) { t.deliver(syn, s, d); };
template <class T, class S, class D>
concept deliver_from_to_stateless = requires(T const t, typename S::neuron const& s, typename D::neuron& d) {
	t.deliver(s, d);
};
}

template <class T, class Neur>
concept DeliverTo = StatefulNeuron<Neur> && ((StatefulSynapse<T> && detail::deliver_to_stateful<T, Neur>) ||
                                             (!StatefulSynapse<T> && detail::deliver_to_stateless<T, Neur>));

template <class T, class SrcNeur, class DstNeur>
concept DeliverFromTo = StatefulNeuron<SrcNeur> && StatefulNeuron<DstNeur> &&
                        ((StatefulSynapse<T> && detail::deliver_from_to_stateful<T, SrcNeur, DstNeur>) ||
                         (!StatefulSynapse<T> && detail::deliver_from_to_stateless<T, SrcNeur, DstNeur>));

template <class T>
concept PlasticSynapse = StatefulSynapse<T> && requires(T const t, typename T::synapse& syn, float dt, bool pre, bool post, Int n) {
	t.update(syn, dt, pre, post);
	t.skip(syn, dt, n);
};

template <class T>
concept PerSynapseInit = StatefulSynapse<T> && requires(T const t, typename T::synapse& syn, Int src, Int dst, std::mt19937& rng) {
	t.init(syn, src, dst, rng);
};

namespace detail {
// DeliverFromTo is only meaningful (and only instantiable) when the source is stateful
template <class T, class S, class D>
constexpr bool deliver_from_to_v = [] {
	if constexpr (StatefulNeuron<S>)
		return DeliverFromTo<T, S, D>;
	else
		return false;
}();
}

template <class T, class SrcNeur, class DstNeur>
concept Synapse = (StatelessSynapse<T> || StatefulSynapse<T> || PlasticSynapse<T>) &&
                  (int(DeliverTo<T, DstNeur>) + int(detail::deliver_from_to_v<T, SrcNeur, DstNeur>) == 1);

namespace detail {
template <class T>
struct neuron_traits {
	using type = void;
};
template <StatefulNeuron T>
struct neuron_traits<T> {
	using type = typename T::neuron;
};
template <class T>
using neuron_traits_t = typename neuron_traits<T>::type;

template <class T>
struct synapse_traits {
	using type = void;
};
template <StatefulSynapse T>
struct synapse_traits<T> {
	using type = typename T::synapse;
};
template <class T>
using synapse_traits_t = typename synapse_traits<T>::type;

// engine draws per update() call; 0 unless the model declares `rng_draws`
template <class T>
constexpr int rng_draws_v = [] {
	if constexpr (requires { T::rng_draws; })
		return int(T::rng_draws);
	else
		return 0;
}();

struct any_neuron_t {
	using neuron = util::any_t;
};
}

// ---- diagnostics (reference: concepts.h:163-228) ----------------------------------------------------
template <class T>
constexpr bool CheckNeuron() {
	constexpr bool has_update = requires(T t, util::any_t a) { t.update(a, a); } ||
	                            requires(T t, util::any_t a) { t.update(a, a, a); };
	static_assert(StatelessNeuron<T>, "Every neuron must at least conform to the StatelessNeuron concept.");
	static_assert(has_update, "Every neuron must define an update() method.");
	static_assert(!has_update || PerNeuronUpdate<T> || PerPopulationUpdate<T>,
	              "Your update() method has the wrong signautre.");
	static_assert(!PerPopulationUpdate<T> ||
	                  util::count_true<StatefulNeuron<T>, PerNeuronUpdate<T>, PerNeuronInit<T>, PerPopulationInit<T>> == 0,
	              "Defining a per-population update() method prohibits you from defining any of the following: "
	              "per-neuron update(), per-neuron init(), per-population init(), making your neuron stateful.");
	constexpr bool has_init = requires(T t, util::any_t a) { t.init(a, a); } ||
	                          requires(T t, util::any_t a) { t.init(a, a, a); };
	static_assert(!has_init || StatefulNeuron<T>, "You defined an init() method but your neuron has no state.");
	static_assert(!has_init || PerNeuronInit<T> || PerPopulationInit<T>,
	              "Your neuron's init() method has the wrong signature.");
	static_assert(util::count_true<PerNeuronInit<T>, PerPopulationInit<T>> <= 1,
	              "Your neuron must define at most 1 init() method.");
	return true;
}

template <class T>
constexpr bool CheckSynapse() {
	constexpr bool has_deliver = requires(T t, util::any_t a) { t.deliver(a); } ||
	                             requires(T t, util::any_t a) { t.deliver(a, a); } ||
	                             requires(T t, util::any_t a) { t.deliver(a, a, a); };
	static_assert(StatelessSynapse<T>, "Every synapse must at least conform to the StatelessSynapse concept.");
	static_assert(has_deliver, "Every synapse must define a deliver() method.");
	static_assert(!has_deliver || DeliverTo<T, detail::any_neuron_t> ||
	                  DeliverFromTo<T, detail::any_neuron_t, detail::any_neuron_t>,
	              "Your deliver() method has the wrong signature.");
	constexpr bool has_update = requires(T t, util::any_t a) { t.update(a, a, a, a); };
	constexpr bool has_skip   = requires(T t, util::any_t a) { t.skip(a, a, a); };
	static_assert(!has_update || StatefulSynapse<T>, "You defined an update() method but your synapse has no state.");
	static_assert(!has_update || has_skip, "Your synapse defines an update() method but no skip() method.");
	static_assert(!has_skip || has_update, "Your synapse defines a skip() method but no update() method.");
	static_assert(!has_update || PlasticSynapse<T>,
	              "Your synapse defines an update() method, suggesting you intend to write a plastic synapse. "
	              "But your synapse does not conform to the PlasticSynapse concept. "
	              "Probably your update() or skip() method have the wrong signature.");
	constexpr bool has_init = requires(T t, util::any_t a) { t.init(a, a, a, a); };
	static_assert(!has_init || StatefulSynapse<T>, "You defined an init() method but your synapse has no state.");
	static_assert(!has_init || PerSynapseInit<T>, "Your synapse's init() method has the wrong signature.");
	return true;
}
}
