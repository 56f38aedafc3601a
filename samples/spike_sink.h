// JSON spike sink for the sample programs: the same stream protocol and byte-for-byte the same
// output as the reference samples' non-plotting sink (samples/matplot.h:12-21,
// samples/matplot.cpp:96-134), so a sample's stdout can be compared with the reference's by md5.
//
//     spike_output_stream s("Brunel");
//     s << I << E << P << '\n';     // once per step, populations in display order
#pragma once

#include <iostream>
#include <string>

#include "spice/snn.h"

inline void pause(double) {}

class spike_output_stream {
public:
	explicit spike_output_stream(std::string const& model_name, bool const skip_steps_without_spikes = false) :
	_skip(skip_steps_without_spikes) {
		std::cout << "{\n\t\"name\": \"" << model_name << "\",\n\t\"spikes\": [\n";
	}
	~spike_output_stream() { std::cout << "\n\t]\n}\n"; }

	// ids are printed relative to the populations streamed so far in this row; a population only
	// advances that base when it is printed (in skip mode an idle population does not)
	spike_output_stream& operator<<(spice::detail::NeuronPopulation const* population) {
		auto const spikes = population->spikes(0);
		if (_skip && spikes.empty())
			return *this;
		for (Int32 const id : spikes) {
			std::cout << _separator << id + _base;
			_separator = ",";
		}
		_base += population->size();
		return *this;
	}

	// '\n' closes the row
	spike_output_stream& operator<<(char const c) {
		if (c == '\n') {
			if (_base > 0) {
				_separator = ",\n\t\t[";
				std::cout << "]";
			}
			_base = 0;
		}
		return *this;
	}

private:
	bool _skip;
	Int _base              = 0;
	std::string _separator = "\t\t[";
};
